"""Torch-facing wrappers over the C ABI.  Inputs must be CUDA fp32 tensors; outputs are
allocated here with torch (caller-owned in ABI terms) and raw pointers + the current stream
are handed to the library."""
import torch

from . import _lib
from ._lib import C2MError, DcnShape

_ws_cache = {}


def _require_cuda(name, t, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f'{name} must be a Tensor')
    if not t.is_cuda:
        # reference: AT_ERROR("Not implemented on the CPU") — DCNv2/src/dcn_v2.h:38
        raise RuntimeError(f'{name} must be a CUDA tensor (Not implemented on the CPU)')
    if t.dtype != dtype:
        raise RuntimeError(f'{name} must be {dtype}, got {t.dtype}')


def _stream():
    return torch.cuda.current_stream().cuda_stream


def launch_count():
    return int(_lib.lib().c2m_launch_count())


def profile_enable(on=True):
    _lib.check(_lib.lib().c2m_profile_enable(int(bool(on))), 'c2m_profile_enable')


PROF_KERNELS = {'corr_search': 0, 'conv3x3': 1, 'dcn': 2}


def profile_collect(kernel):
    """{'ms', 'launches', 'flops', 'bytes'} of one kernel class since the last collect."""
    import ctypes
    ms, n, fl, by = ctypes.c_float(0), ctypes.c_int(0), ctypes.c_double(0), ctypes.c_double(0)
    _lib.check(_lib.lib().c2m_profile_collect(PROF_KERNELS[kernel], ctypes.byref(ms), ctypes.byref(n),
                                              ctypes.byref(fl), ctypes.byref(by)), 'c2m_profile_collect')
    return {'ms': float(ms.value), 'launches': int(n.value), 'flops': float(fl.value), 'bytes': float(by.value)}


def _workspace(nbytes, device):
    """Per-(device, stream) scratch, grown on demand.  Re-entrant across DataParallel threads:
    each replica runs on its own device."""
    key = (device.index, _stream())
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def corr_argmax(feat_in, feat_ref, patch_size=3, input_stride=1, ref_stride=1, is_norm=True,
                norm_input=False, l2norm=False, force_generic=False):
    """Batched correlation + argmax.  feat_in [B,C,h,w], feat_ref [B,C,hr,wr] ->
    (idx int64 [B,h',w'], val fp32 [B,h',w']).  See include/c2m_sm100.h."""
    _require_cuda('feat_input', feat_in)
    _require_cuda('feat_ref', feat_ref)
    if feat_in.dim() != 4 or feat_ref.dim() != 4:
        raise RuntimeError('corr_argmax expects [B,C,h,w] tensors')
    if feat_in.shape[:2] != feat_ref.shape[:2]:
        raise RuntimeError(f'batch/channel mismatch: {tuple(feat_in.shape)} vs {tuple(feat_ref.shape)}')
    feat_in = feat_in.contiguous()
    feat_ref = feat_ref.contiguous()
    B, C, h, w = feat_in.shape
    hr, wr = feat_ref.shape[2:]
    L = _lib.lib()
    with torch.cuda.device(feat_in.device):
        need = L.c2m_corr_workspace_bytes(B, C, h, w, hr, wr, patch_size, input_stride, ref_stride)
        if need == 0:
            raise C2MError('corr_argmax: ' + L.c2m_last_error().decode())
        ws = _workspace(need, feat_in.device)
        gh, gw = (h - patch_size) // input_stride + 1, (w - patch_size) // input_stride + 1
        idx = torch.empty(B, gh, gw, dtype=torch.int64, device=feat_in.device)
        val = torch.empty(B, gh, gw, dtype=torch.float32, device=feat_in.device)
        rc = L.c2m_corr_argmax_f32(feat_in.data_ptr(), feat_ref.data_ptr(), B, C, h, w, hr, wr, patch_size,
                                   input_stride, ref_stride, int(is_norm), int(norm_input), int(l2norm),
                                   1 if force_generic else 0, idx.data_ptr(), val.data_ptr(), ws.data_ptr(),
                                   ws.numel(), _stream())
        _lib.check(rc, 'c2m_corr_argmax_f32')
    return idx, val


def feature_match_index(feat_input, feat_ref, patch_size=3, input_stride=1, ref_stride=1, is_norm=True,
                        norm_input=False):
    """Drop-in for mmsr/models/archs/ref_map_util.py:26-86 (same signature, same return)."""
    idx, val = corr_argmax(feat_input.unsqueeze(0), feat_ref.unsqueeze(0), patch_size, input_stride, ref_stride,
                           is_norm, norm_input)
    return idx[0], val[0]


def offset_pyramid(idx, scale, ref_gw=None):
    """idx int64 [B,gh,gw] -> pre_offset [B,9,s*(gh+2),s*(gw+2),2] (x,y)."""
    _require_cuda('idx', idx, torch.int64)
    idx = idx.contiguous()
    B, gh, gw = idx.shape
    out = torch.empty(B, 9, scale * (gh + 2), scale * (gw + 2), 2, dtype=torch.float32, device=idx.device)
    with torch.cuda.device(idx.device):
        rc = _lib.lib().c2m_offset_pyramid_f32(idx.data_ptr(), B, gh, gw, ref_gw or gw, scale, out.data_ptr(),
                                               _stream())
        _lib.check(rc, 'c2m_offset_pyramid_f32')
    return out


def _dense_nchw_or_nhwc(t):
    """Accept NCHW-contiguous or channels-last storage without copying; otherwise make NCHW."""
    if t.is_contiguous() or t.is_contiguous(memory_format=torch.channels_last):
        return t
    return t.contiguous()


def _shape(x, cout, kh, kw, sh, sw, ph, pw, dh, dw, dg, out):
    s = DcnShape()
    s.B, s.C, s.H, s.W = x.shape
    s.Cout = cout
    s.kh, s.kw, s.sh, s.sw, s.ph, s.pw, s.dh, s.dw, s.dg = kh, kw, sh, sw, ph, pw, dh, dw, dg
    s.xs_b, s.xs_c, s.xs_y, s.xs_x = x.stride()
    s.os_b, s.os_c, s.os_y, s.os_x = out.stride()
    return s


def _out_hw(H, W, kh, kw, sh, sw, ph, pw, dh, dw):
    return ((H + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1, (W + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1)


def dcn_v2_forward(input, weight, bias, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                   dilation_h, dilation_w, deformable_group):
    """`_ext.dcn_v2_forward` (DCNv2/src/dcn_v2.h:9-39; call site dcn_v2.py:26-30)."""
    for n, t in (('input', input), ('weight', weight), ('bias', bias), ('offset', offset), ('mask', mask)):
        _require_cuda(n, t)
    if weight.shape[2] != kernel_h or weight.shape[3] != kernel_w:
        raise RuntimeError(f'Input shape and kernel shape wont match: ({kernel_h} x {kernel_w} vs '
                           f'{weight.shape[2]} x {weight.shape[3]}).')
    if input.shape[1] != weight.shape[1]:
        raise RuntimeError(f'Input shape and kernel channels wont match: ({input.shape[1]} vs {weight.shape[1]}).')
    x = _dense_nchw_or_nhwc(input)
    B, C, H, W = x.shape
    Ho, Wo = _out_hw(H, W, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w)
    T = kernel_h * kernel_w
    if tuple(offset.shape) != (B, 2 * deformable_group * T, Ho, Wo) or tuple(mask.shape) != (B, deformable_group * T, Ho, Wo):
        raise RuntimeError(f'offset/mask shape mismatch: {tuple(offset.shape)}, {tuple(mask.shape)} for output {Ho}x{Wo}')
    import os
    if ((kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w) == (3, 3, 1, 1, 1, 1, 1, 1) and
            dcn_tc_supported(C, weight.shape[0], deformable_group) and os.environ.get('C2M_EXT_DCN_TC', '1') != '0'):
        # eligible shapes (every DCN of C2-Matching) take the tensor-core kernel: fp32 NCHW -> packed-split operand
        # in one pass, final offsets / mask read as given, fp32 NCHW out.  The FFMA kernel below stays the general
        # path (any kernel size / stride / dilation / group width).
        return dcn_v2_fused_tc(x, offset.contiguous(), weight, bias, deformable_group, final_mask=mask.contiguous())
    out = torch.empty(B, weight.shape[0], Ho, Wo, dtype=torch.float32, device=x.device)
    s = _shape(x, weight.shape[0], kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
               deformable_group, out)
    with torch.cuda.device(x.device):
        rc = _lib.lib().c2m_dcn_v2_forward_f32(x.data_ptr(), offset.contiguous().data_ptr(),
                                               mask.contiguous().data_ptr(), weight.contiguous().data_ptr(),
                                               bias.contiguous().data_ptr(), out.data_ptr(), s, _stream())
        _lib.check(rc, 'c2m_dcn_v2_forward_f32')
    return out


def dcn_v2_fused_forward(x, om, weight, bias, deformable_group, pre_offset=None, idx=None, pre_scale=1,
                         ref_gw=None, lrelu_slope=1.0, channels_last_out=False):
    """Tail of DCN_sep_pre_multi_offset.forward (dcn_v2.py:229-253) in one kernel: `om` is the raw
    conv_offset_mask output.  Either `pre_offset` [B,9,H,W,2] or `idx` [B,gh,gw] (+pre_scale)."""
    _require_cuda('x', x)
    _require_cuda('om', om)
    x = _dense_nchw_or_nhwc(x)
    om = om.contiguous()
    B, C, H, W = x.shape
    kh, kw = weight.shape[2:]
    Ho, Wo = _out_hw(H, W, kh, kw, 1, 1, kh // 2, kw // 2, 1, 1)
    if tuple(om.shape) != (B, 3 * deformable_group * kh * kw, Ho, Wo):
        raise RuntimeError(f'conv_offset_mask output has shape {tuple(om.shape)}')
    mf = torch.channels_last if channels_last_out else torch.contiguous_format
    out = torch.empty(B, weight.shape[0], Ho, Wo, dtype=torch.float32, device=x.device, memory_format=mf)
    s = _shape(x, weight.shape[0], kh, kw, 1, 1, kh // 2, kw // 2, 1, 1, deformable_group, out)
    pre_p = idx_p = None
    gh = gw = 0
    if pre_offset is not None:
        _require_cuda('pre_offset', pre_offset)
        if tuple(pre_offset.shape) != (B, kh * kw, Ho, Wo, 2):
            raise RuntimeError(f'pre_offset has shape {tuple(pre_offset.shape)}')
        pre_offset = pre_offset.contiguous()
        pre_p = pre_offset.data_ptr()
    elif idx is not None:
        _require_cuda('idx', idx, torch.int64)
        idx = idx.contiguous()
        gh, gw = idx.shape[1:]
        idx_p = idx.data_ptr()
    with torch.cuda.device(x.device):
        rc = _lib.lib().c2m_dcn_v2_fused_forward_f32(
            x.data_ptr(), om.data_ptr(), pre_p, idx_p, gh, gw, ref_gw or gw, pre_scale,
            weight.contiguous().data_ptr(), bias.contiguous().data_ptr() if bias is not None else None,
            float(lrelu_slope), out.data_ptr(), s, _stream())
        _lib.check(rc, 'c2m_dcn_v2_fused_forward_f32')
    return out


# ------------------------------------------------------------------------------------------------
# Packed-split activations ("PSA") and the tcgen05 3x3 convolution (include/c2m_sm100.h)
class PSA:
    """fp16 hi/lo planes [B][ceil(C/8)][H][W][8]; value = (hi + lo) * 2^-sa."""
    __slots__ = ('hi', 'lo', 'B', 'C', 'H', 'W', 'sa')

    def __init__(self, hi, lo, B, C, H, W, sa=0):
        self.hi, self.lo, self.B, self.C, self.H, self.W, self.sa = hi, lo, B, C, H, W, sa

    @staticmethod
    def empty(B, C, H, W, device, sa=0):
        c8 = (C + 7) // 8
        buf = torch.empty(2, B, c8, H, W, 8, dtype=torch.float16, device=device)
        return PSA(buf[0], buf[1], B, C, H, W, sa)

    @property
    def shape(self):
        return (self.B, self.C, self.H, self.W)


class OctF32:
    """fp32 tensor in the octet-planar layout [B][ceil(C/8)][H][W][8] (padding channels are 0): what
    conv3x3_psa(..., f32_octets=True) writes with two 16 B stores per octet and what the tensor-core DCN reads
    its offsets / mask from.  `.nchw()` gives the ordinary [B,C,H,W] tensor."""
    __slots__ = ('data', 'C')

    def __init__(self, data, C):
        self.data, self.C = data, C

    @property
    def shape(self):
        B, _, H, W, _ = self.data.shape
        return (B, self.C, H, W)

    def nchw(self):
        B, c8, H, W, _ = self.data.shape
        return self.data.permute(0, 1, 4, 2, 3).reshape(B, c8 * 8, H, W)[:, :self.C].contiguous()


def suggest_sa(amax, target_log2=12):
    """Scale exponent for a PSA tensor whose largest magnitude is `amax`: amax * 2^sa ~ 2^target_log2 (fp16 tops out
    at 65504 = 2^16, so 16x headroom; values below 2^-14 * 2^-sa lose their lo half to fp16 subnormals)."""
    import math
    if not (amax > 0) or not math.isfinite(amax):
        return 0
    return max(-14, min(24, target_log2 - math.ceil(math.log2(amax))))


def psa_amax(p):
    """Largest |value| of a PSA tensor (from its hi half: within 2^-11), as a Python float (synchronises)."""
    return float(p.hi.abs().max()) * 2.0 ** (-p.sa)


def psa_from_f32(x, sa=0):
    _require_cuda('x', x)
    B, C, H, W = x.shape
    out = PSA.empty(B, C, H, W, x.device, sa)
    with torch.cuda.device(x.device):
        rc = _lib.lib().c2m_psa_from_f32(x.data_ptr(), B, C, H, W, *x.stride(), sa, out.hi.data_ptr(),
                                         out.lo.data_ptr(), _stream())
        _lib.check(rc, 'c2m_psa_from_f32')
    return out


def psa_to_f32(p, add=None, channels_last=False):
    mf = torch.channels_last if channels_last else torch.contiguous_format
    out = torch.empty(p.B, p.C, p.H, p.W, dtype=torch.float32, device=p.hi.device, memory_format=mf)
    if add is not None:
        _require_cuda('add', add)
        if add.shape != out.shape or add.stride() != out.stride():
            add = add.contiguous(memory_format=mf)
    with torch.cuda.device(out.device):
        rc = _lib.lib().c2m_psa_to_f32(p.hi.data_ptr(), p.lo.data_ptr(), p.B, p.C, p.H, p.W, p.sa,
                                       add.data_ptr() if add is not None else None, out.data_ptr(), *out.stride(),
                                       _stream())
        _lib.check(rc, 'c2m_psa_to_f32')
    return out


def psa_interleave(p):
    """PSA planes -> fp16 [B][ceil(C/8)][H][W][16] with the hi and lo octets of a pixel adjacent (the tensor-core DCN's
    gather operand: one 32 B sector per corner fetch instead of two)."""
    if p.sa != 0:
        raise RuntimeError('psa_interleave: scale exponent must be 0')
    out = torch.empty(p.B, (p.C + 7) // 8, p.H, p.W, 16, dtype=torch.float16, device=p.hi.device)
    with torch.cuda.device(p.hi.device):
        rc = _lib.lib().c2m_psa_interleave(p.hi.data_ptr(), p.lo.data_ptr(), p.B, p.C, p.H, p.W, out.data_ptr(), _stream())
        _lib.check(rc, 'c2m_psa_interleave')
    return out


def psa_maxpool2(p):
    """nn.MaxPool2d(kernel_size=2, stride=2) on a PSA tensor."""
    out = PSA.empty(p.B, p.C, p.H // 2, p.W // 2, p.hi.device, p.sa)
    with torch.cuda.device(p.hi.device):
        rc = _lib.lib().c2m_psa_maxpool2(p.hi.data_ptr(), p.lo.data_ptr(), p.B, p.C, p.H, p.W, out.hi.data_ptr(),
                                         out.lo.data_ptr(), _stream())
        _lib.check(rc, 'c2m_psa_maxpool2')
    return out


def conv3x3_supported(cin, cout, H=None, W=None):
    """The tcgen05 kernel takes any channel counts; maps must hold one halo tile (18 x 10)."""
    return H is None or (H >= 18 and W >= 10)


# ---- packed-weight cache.  The blob lives ON the tensor object (so it dies with it — a global cache keyed
# by data_ptr would alias a freed parameter's address) and is re-made when the storage, the shape or the
# version counter changes.  Writes through `.data` do NOT bump the version: call invalidate_packs() after
# such an edit (load_state_dict / `with torch.no_grad(): p.copy_()` do bump it).
_fallback_packs = {}          # id(tensor) -> (weakref, {attr: entry}) for tensors that refuse attributes


def _pack_cached(weight, attr, extra, make):
    import weakref
    tag = (weight.data_ptr(), weight._version, tuple(weight.shape), extra)
    store = getattr(weight, '__dict__', None)
    if store is None:
        ent = _fallback_packs.get(id(weight))
        if ent is None or ent[0]() is not weight:
            key = id(weight)
            ent = (weakref.ref(weight, lambda _r, key=key: _fallback_packs.pop(key, None)), {})
            _fallback_packs[key] = ent
        store = ent[1]
    cached = store.get(attr)
    cur = torch.cuda.current_stream(weight.device)
    if cached is not None and cached[0] == tag:
        _, blob, ev, st = cached
        if st != cur.cuda_stream and not torch.cuda.is_current_stream_capturing():
            # packed on another stream: order this stream after the pack kernels.  (A capturing stream cannot wait on
            # an event recorded outside its capture; RestorationPipeline synchronises the device before it captures.)
            cur.wait_event(ev)
        return blob
    blob = make()
    ev = torch.cuda.Event()
    ev.record(cur)
    store[attr] = (tag, blob, ev, cur.cuda_stream)
    return blob


def invalidate_packs(obj):
    """Drop the cached packed weights of a tensor, or of every parameter of a module (needed after in-place
    edits through `.data`, which do not bump the tensor version)."""
    tensors = [obj] if isinstance(obj, torch.Tensor) else list(obj.parameters())
    if not isinstance(obj, torch.Tensor):
        for m in obj.modules():                 # calibrated PSA scale exponents of VGG trunks depend on the weights too
            m.__dict__.pop('_c2m_sa', None)
    for t in tensors:
        d = getattr(t, '__dict__', None)
        if d is not None:
            d.pop('_c2m_pack', None)
            d.pop('_c2m_dcn_pack', None)
        _fallback_packs.pop(id(t), None)


def conv3x3_pack_weights(weight):
    """Pack a [Cout,Cin,3,3] fp32 weight for c2m_conv3x3_psa (cached, see _pack_cached)."""
    _require_cuda('weight', weight)
    cout, cin = weight.shape[:2]

    def make():
        n = _lib.lib().c2m_conv3x3_packed_weight_bytes(cin, cout)
        blob = torch.empty(n, dtype=torch.uint8, device=weight.device)
        with torch.cuda.device(weight.device):
            rc = _lib.lib().c2m_conv3x3_pack_weights_f32(weight.detach().contiguous().data_ptr(), cin, cout,
                                                         blob.data_ptr(), _stream())
            _lib.check(rc, 'c2m_conv3x3_pack_weights_f32')
        return blob
    return _pack_cached(weight, '_c2m_pack', None, make)


_ACT = {None: 0, 'none': 0, 'relu': 1, 'lrelu': 2}


def conv3x3_psa(x, weight, bias, act=None, residual=None, residual2=None, x2=None, out=None, sa_out=0,
                pixel_shuffle=0, out_f32=False, add_f32=None, psa_out=True, channels_last=False, f32_octets=False):
    """y = act(conv3x3(cat[x, x2], weight) + bias), fp32-grade on tensor cores.
    f32_octets: the fp32 result is an OctF32 (octet-planar) instead of a strided tensor.

    Returns the PSA tensor `y + residual + residual2` (or PixelShuffle(2)(y)), and/or — with
    out_f32=True — the fp32 tensor `y + add_f32`.  With both, returns (psa, f32)."""
    cout, cin = weight.shape[:2]
    c_in = x.C + (x2.C if x2 is not None else 0)
    if cin != c_in:
        raise RuntimeError(f'conv3x3_psa: weight expects {cin} input channels, got {c_in}')
    blob = conv3x3_pack_weights(weight)
    dev = x.hi.device
    a = _lib.ConvArgs()
    a.in_hi, a.in_lo, a.Cin = x.hi.data_ptr(), x.lo.data_ptr(), x.C
    if x2 is not None:
        if (x2.B, x2.H, x2.W) != (x.B, x.H, x.W) or x2.sa != x.sa:
            raise RuntimeError('conv3x3_psa: concatenated inputs must share batch / size / scale')
        a.in2_hi, a.in2_lo, a.Cin2 = x2.hi.data_ptr(), x2.lo.data_ptr(), x2.C
    a.B, a.H, a.W, a.sa_in = x.B, x.H, x.W, x.sa
    a.packed_w, a.bias = blob.data_ptr(), (bias.data_ptr() if bias is not None else None)
    a.Cout, a.act = cout, _ACT[act]
    for r, (fh, fl) in ((residual, ('res_hi', 'res_lo')), (residual2, ('res2_hi', 'res2_lo'))):
        if r is not None:
            if r.shape != (x.B, cout, x.H, x.W):
                raise RuntimeError(f'conv3x3_psa: residual shape {r.shape}')
            setattr(a, fh, r.hi.data_ptr())
            setattr(a, fl, r.lo.data_ptr())
            a.sa_res = r.sa
    a.pixel_shuffle = pixel_shuffle
    ret_psa = None
    if psa_out:
        if out is None:
            if pixel_shuffle == 2:
                out = PSA.empty(x.B, cout // 4, 2 * x.H, 2 * x.W, dev, sa_out)
            else:
                out = PSA.empty(x.B, cout, x.H, x.W, dev, sa_out)
        a.out_hi, a.out_lo, a.sa_out = out.hi.data_ptr(), out.lo.data_ptr(), out.sa
        ret_psa = out
    ret_f32 = None
    if out_f32 and f32_octets:
        if add_f32 is not None:
            raise RuntimeError('conv3x3_psa: add_f32 is not available with f32_octets')
        ret_f32 = OctF32(torch.empty(x.B, (cout + 7) // 8, x.H, x.W, 8, dtype=torch.float32, device=dev), cout)
        a.out_f32, a.out_f32_octets = ret_f32.data.data_ptr(), 1
    elif out_f32:
        mf = torch.channels_last if channels_last else torch.contiguous_format
        ret_f32 = torch.empty(x.B, cout, x.H, x.W, dtype=torch.float32, device=dev, memory_format=mf)
        a.out_f32 = ret_f32.data_ptr()
        a.os_b, a.os_c, a.os_y, a.os_x = ret_f32.stride()
        if add_f32 is not None:
            if add_f32.shape != ret_f32.shape or add_f32.stride() != ret_f32.stride():
                add_f32 = add_f32.contiguous(memory_format=mf)
            a.add_f32 = add_f32.data_ptr()
    import ctypes
    with torch.cuda.device(dev):
        rc = _lib.lib().c2m_conv3x3(ctypes.addressof(a), _stream())
        _lib.check(rc, 'c2m_conv3x3')
    if ret_psa is not None and ret_f32 is not None:
        return ret_psa, ret_f32
    return ret_psa if ret_psa is not None else ret_f32


# ------------------------------------------------------------------------------------------------
# DCNv2 on tensor cores
def dcn_tc_supported(C, Cout, dg, kh=3, kw=3):
    return kh == 3 and kw == 3 and bool(_lib.lib().c2m_dcn_tc_supported(C, Cout, dg))


def _dcn_tc_pack(weight, dg):
    cout, c = weight.shape[:2]

    def make():
        n = _lib.lib().c2m_dcn_tc_packed_weight_bytes(c, cout, dg)
        blob = torch.empty(n, dtype=torch.uint8, device=weight.device)
        with torch.cuda.device(weight.device):
            rc = _lib.lib().c2m_dcn_tc_pack_weights_f32(weight.detach().contiguous().data_ptr(), c, cout, dg,
                                                        blob.data_ptr(), _stream())
            _lib.check(rc, 'c2m_dcn_tc_pack_weights_f32')
        return blob
    return _pack_cached(weight, '_c2m_dcn_pack', dg, make)


def dcn_v2_fused_tc(x, om, weight, bias, deformable_group, pre_offset=None, idx=None, pre_scale=1, ref_gw=None,
                    lrelu=False, psa_out=False, out_f32=True, channels_last_out=False, final_mask=None):
    """Tensor-core version of dcn_v2_fused_forward (3x3/s1/p1/d1).  Returns fp32 and / or PSA.
    final_mask: `om` then holds the FINAL offsets [B,2*dg*9,H,W] and final_mask the FINAL modulation
    [B,dg*9,H,W] (the `_ext.dcn_v2_forward` contract: no pre-offsets, no sigmoid)."""
    import ctypes
    om_oct = isinstance(om, OctF32)
    _require_cuda('om', om.data if om_oct else om)
    if isinstance(x, PSA):
        xp = x
    else:
        _require_cuda('x', x)
        xp = getattr(x, '_c2m_psa', None)
        if xp is None or xp.shape != tuple(x.shape):
            xp = psa_from_f32(x)
    if xp.sa != 0:
        raise RuntimeError('dcn_v2_fused_tc: PSA input must have scale exponent 0')
    B, C, H, W = xp.shape
    dev = xp.hi.device
    cout = weight.shape[0]
    if final_mask is not None:
        _require_cuda('final_mask', final_mask)
        if om_oct or pre_offset is not None or idx is not None:
            raise RuntimeError('dcn_v2_fused_tc: final_mask excludes octet-planar om and pre-offsets')
        if tuple(om.shape) != (B, 18 * deformable_group, H, W) or tuple(final_mask.shape) != (B, 9 * deformable_group, H, W):
            raise RuntimeError(f'offset/mask shape mismatch: {tuple(om.shape)}, {tuple(final_mask.shape)}')
    elif tuple(om.shape) != (B, 27 * deformable_group, H, W):
        raise RuntimeError(f'conv_offset_mask output has shape {tuple(om.shape)}')
    a = _lib.DcnTcArgs()
    import os
    x_il = None
    if os.environ.get('C2M_DCN_INTERLEAVE', '1') != '0':
        x_il = psa_interleave(xp)           # kept alive until the launch below is enqueued (same stream)
        a.x_il = x_il.data_ptr()
    if final_mask is not None:
        final_mask = final_mask.contiguous()
        a.mask = final_mask.data_ptr()
    a.x_hi, a.x_lo = xp.hi.data_ptr(), xp.lo.data_ptr()
    if om_oct:
        a.om, a.om_octets = om.data.data_ptr(), 1
    else:
        om = om.contiguous()
        a.om = om.data_ptr()
    if pre_offset is not None:
        _require_cuda('pre_offset', pre_offset)
        if tuple(pre_offset.shape) != (B, 9, H, W, 2):
            raise RuntimeError(f'pre_offset has shape {tuple(pre_offset.shape)}')
        pre_offset = pre_offset.contiguous()
        a.pre = pre_offset.data_ptr()
    elif idx is not None:
        _require_cuda('idx', idx, torch.int64)
        idx = idx.contiguous()
        a.idx = idx.data_ptr()
        a.gh, a.gw = idx.shape[1:]
        a.ref_gw, a.pre_scale = (ref_gw or idx.shape[2]), pre_scale
    a.B, a.C, a.H, a.W, a.Cout, a.dg = B, C, H, W, cout, deformable_group
    blob = _dcn_tc_pack(weight, deformable_group)
    a.packed_w = blob.data_ptr()
    a.bias = bias.contiguous().data_ptr() if bias is not None else None
    a.lrelu = int(bool(lrelu))
    ret_p = ret_f = None
    if psa_out:
        ret_p = PSA.empty(B, cout, H, W, dev)
        a.out_hi, a.out_lo, a.sa_out = ret_p.hi.data_ptr(), ret_p.lo.data_ptr(), 0
    if out_f32:
        mf = torch.channels_last if channels_last_out else torch.contiguous_format
        ret_f = torch.empty(B, cout, H, W, dtype=torch.float32, device=dev, memory_format=mf)
        a.out_f32 = ret_f.data_ptr()
        a.os_b, a.os_c, a.os_y, a.os_x = ret_f.stride()
    with torch.cuda.device(dev):
        rc = _lib.lib().c2m_dcn_v2_fused_tc(ctypes.addressof(a), _stream())
        _lib.check(rc, 'c2m_dcn_v2_fused_tc')
    if ret_p is not None and ret_f is not None:
        return ret_p, ret_f
    return ret_p if ret_p is not None else ret_f


# ------------------------------------------------------------------------------------------------
# DCNv2 backward (training; completes the `_ext` ABI)
def dcn_v2_backward(input, weight, bias, offset, mask, grad_output, kernel_h, kernel_w, stride_h, stride_w, pad_h,
                    pad_w, dilation_h, dilation_w, deformable_group):
    """`_ext.dcn_v2_backward` (DCNv2/src/dcn_v2.h:41-72; call site dcn_v2.py:38-47) ->
    [grad_input, grad_offset, grad_mask, grad_weight, grad_bias].  The deformable pieces
    (im2col, coordinate/mask gradient, input scatter) are libc2m_sm100 kernels; the two dense
    GEMMs are torch.matmul in full fp32."""
    for n, t in (('input', input), ('weight', weight), ('bias', bias), ('offset', offset), ('mask', mask),
                 ('grad_output', grad_output)):
        _require_cuda(n, t)
    x, w = input.contiguous(), weight.contiguous()
    offset, mask, gout = offset.contiguous(), mask.contiguous(), grad_output.contiguous()
    B, C, H, W = x.shape
    cout = w.shape[0]
    Ho, Wo = _out_hw(H, W, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w)
    T, P = kernel_h * kernel_w, Ho * Wo
    if tuple(gout.shape) != (B, cout, Ho, Wo):
        raise RuntimeError(f'grad_output has shape {tuple(gout.shape)}, expected {(B, cout, Ho, Wo)}')
    s = _shape(x, cout, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w, deformable_group,
               gout)
    L = _lib.lib()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.cuda.device(x.device):
            w2 = w.view(cout, C * T)
            g2 = gout.view(B, cout, P)
            gcol = torch.matmul(w2.t(), g2).contiguous()                          # [B, C*T, P]
            grad_offset, grad_mask = torch.empty_like(offset), torch.empty_like(mask)
            _lib.check(L.c2m_dcn_v2_col2im_coord_f32(gcol.data_ptr(), x.data_ptr(), offset.data_ptr(), mask.data_ptr(),
                                                     s, grad_offset.data_ptr(), grad_mask.data_ptr(), _stream()),
                       'c2m_dcn_v2_col2im_coord_f32')
            grad_input = torch.zeros_like(x)
            _lib.check(L.c2m_dcn_v2_col2im_f32(gcol.data_ptr(), offset.data_ptr(), mask.data_ptr(), s,
                                               grad_input.data_ptr(), _stream()), 'c2m_dcn_v2_col2im_f32')
            columns = gcol                                                        # reuse the buffer
            _lib.check(L.c2m_dcn_v2_im2col_f32(x.data_ptr(), offset.data_ptr(), mask.data_ptr(), s, columns.data_ptr(),
                                               _stream()), 'c2m_dcn_v2_im2col_f32')
            grad_weight = torch.matmul(g2, columns.transpose(1, 2)).sum(0).view_as(w)
            grad_bias = gout.sum(dim=(0, 2, 3))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return [grad_input, grad_offset, grad_mask, grad_weight, grad_bias]
