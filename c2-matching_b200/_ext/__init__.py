"""Drop-in replacement for the reference's compiled top-level module `_ext`
(mmsr/models/archs/DCNv2/src/vision.cpp:3-8; imported as `import _ext as _backend` at
mmsr/models/archs/DCNv2/dcn_v2.py:6).  Same names, positional signatures and error behaviour;
the arithmetic runs in libc2m_sm100.so through the C ABI (include/c2m_sm100.h)."""
from c2m_b200.ops import dcn_v2_forward  # noqa: F401


def dcn_v2_backward(input, weight, bias, offset, mask, grad_output, kernel_h, kernel_w, stride_h, stride_w,
                    pad_h, pad_w, dilation_h, dilation_w, deformable_group):
    """DCNv2/src/dcn_v2.h:41-72.  Training is outside the restoration-forward hot path
    (SURVEY.md §8f N4); fail loudly rather than silently compute nothing."""
    raise NotImplementedError('_ext.dcn_v2_backward: the B200 build covers the inference forward only')


def dcn_v2_psroi_pooling_forward(*args, **kwargs):
    """DCNv2/src/dcn_v2.h:75-108 — unused by any C2-Matching arch."""
    raise NotImplementedError('_ext.dcn_v2_psroi_pooling_forward is not part of the C2-Matching hot path')


def dcn_v2_psroi_pooling_backward(*args, **kwargs):
    """DCNv2/src/dcn_v2.h:110-144 — unused by any C2-Matching arch."""
    raise NotImplementedError('_ext.dcn_v2_psroi_pooling_backward is not part of the C2-Matching hot path')
