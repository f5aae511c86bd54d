"""Model registry: every `*_model.py` in this folder is imported; `create_model(opt)` looks
`opt['model_type']` up by class name (reference: mmsr/models/__init__.py:8-43)."""
import importlib
import logging
import os

_here = os.path.dirname(os.path.abspath(__file__))
_model_modules = [
    importlib.import_module(f'mmsr.models.{name[:-3]}')
    for name in sorted(os.listdir(_here)) if name.endswith('_model.py')
]


def create_model(opt):
    model_type = opt['model_type']
    for module in _model_modules:
        model_cls = getattr(module, model_type, None)
        if model_cls is not None:
            model = model_cls(opt)
            logging.getLogger('base').info(f'Model [{model.__class__.__name__}] is created.')
            return model
    raise ValueError(f'Model {model_type} is not found.')
