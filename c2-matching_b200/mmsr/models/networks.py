"""yaml `type:` -> class lookup over the scanned arch modules (reference:
mmsr/models/networks.py:4-76).  `opt['network_*']` dicts are consumed exactly as there:
`type` is popped, the rest are constructor kwargs."""
from mmsr.models.archs import _arch_modules


def dynamical_instantiation(modules, cls_type, opt):
    for module in modules:
        cls_ = getattr(module, cls_type, None)
        if cls_ is not None:
            return cls_(**opt)
    raise ValueError(f'{cls_type} is not found.')


def _define(section):
    def define(opt):
        opt_net = opt[section]
        network_type = opt_net.pop('type')
        return dynamical_instantiation(_arch_modules, network_type, opt_net)
    define.__name__ = 'define_' + section.replace('network', 'net')
    return define


define_net_g = _define('network_g')
define_net_d = _define('network_d')
define_net_map = _define('network_map')
define_net_extractor = _define('network_extractor')
define_net_student = _define('network_student')
define_net_teacher = _define('network_teacher')
