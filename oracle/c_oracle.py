"""ORACLE — test infrastructure, NOT product code.

ctypes front-end for oracle/c2m_oracle.c (the literal plain-C restatement).  `build()`
compiles it with gcc into oracle/_build/ (git-ignored; travels to the GPU box with the
snapshot).  Only tests/, __graft_entry__ and bench.py's cpu_baseline leg import this."""
import ctypes
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'c2m_oracle.c')
OUT_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(OUT_DIR, 'libc2m_oracle.so')
_lib = None


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ['gcc', '-O2', '-fopenmp', '-fPIC', '-shared', '-o', LIB, SRC, '-lm']
    subprocess.run(cmd, check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = ctypes.CDLL(LIB)
    return _lib


def _p(t, ty):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ty))


def corr_argmax(fin, fref, patch=3, s_in=1, s_ref=1, is_norm=True, norm_input=False, want_gap=False):
    fin = fin.contiguous().float()
    fref = fref.contiguous().float()
    c, h, w = fin.shape
    _, hr, wr = fref.shape
    nh, nw = (h - patch) // s_in + 1, (w - patch) // s_in + 1
    idx = torch.empty(nh, nw, dtype=torch.int64)
    val = torch.empty(nh, nw, dtype=torch.float32)
    gap = torch.empty(nh, nw, dtype=torch.float64) if want_gap else None
    rc = lib().oracle_corr_argmax(
        _p(fin, ctypes.c_float), _p(fref, ctypes.c_float), c, h, w, hr, wr, patch, s_in, s_ref,
        int(is_norm), int(norm_input), _p(idx, ctypes.c_int64), _p(val, ctypes.c_float),
        _p(gap, ctypes.c_double) if want_gap else None)
    if rc:
        raise RuntimeError(f'oracle_corr_argmax rc={rc}')
    return (idx, val, gap) if want_gap else (idx, val)


def offset_pyramid(idx, s, ref_gw=None):
    idx = idx.contiguous().to(torch.int64)
    gh, gw = idx.shape
    out = torch.empty(9, (gh + 2) * s, (gw + 2) * s, 2, dtype=torch.float32)
    lib().oracle_offset_pyramid(_p(idx, ctypes.c_int64), gh, gw, ref_gw or gw, s, _p(out, ctypes.c_float))
    return out


def dcn_v2_forward(x, weight, bias, offset, mask, kh=3, kw=3, sh=1, sw=1, ph=1, pw=1, dh=1, dw=1,
                   dg=8, acc64=True):
    x, weight, bias, offset, mask = [t.contiguous().float() for t in (x, weight, bias, offset, mask)]
    b, c, h, w = x.shape
    cout = weight.shape[0]
    ho = (h + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    wo = (w + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    out = torch.empty(b, cout, ho, wo, dtype=torch.float32)
    rc = lib().oracle_dcn_v2_forward(
        _p(x, ctypes.c_float), _p(weight, ctypes.c_float), _p(bias, ctypes.c_float),
        _p(offset, ctypes.c_float), _p(mask, ctypes.c_float), _p(out, ctypes.c_float),
        b, c, h, w, cout, kh, kw, sh, sw, ph, pw, dh, dw, dg, int(acc64))
    if rc:
        raise RuntimeError(f'oracle_dcn_v2_forward rc={rc}')
    return out


def channel_l2norm(x):
    x = x.contiguous().float()
    y = torch.empty_like(x)
    lib().oracle_channel_l2norm(_p(x, ctypes.c_float), _p(y, ctypes.c_float), x.shape[0],
                                ctypes.c_long(x[0].numel()))
    return y


def dcn_v2_backward(x, weight, bias, offset, mask, grad_output, kh=3, kw=3, sh=1, sw=1, ph=1, pw=1, dh=1, dw=1, dg=8):
    x, weight, offset, mask, grad_output = [t.contiguous().float() for t in (x, weight, offset, mask, grad_output)]
    b, c, h, w = x.shape
    cout = weight.shape[0]
    gx, goff, gm = torch.empty_like(x), torch.empty_like(offset), torch.empty_like(mask)
    gw, gb = torch.empty_like(weight), torch.empty(cout)
    rc = lib().oracle_dcn_v2_backward(
        _p(x, ctypes.c_float), _p(weight, ctypes.c_float), _p(offset, ctypes.c_float), _p(mask, ctypes.c_float),
        _p(grad_output, ctypes.c_float), _p(gx, ctypes.c_float), _p(goff, ctypes.c_float), _p(gm, ctypes.c_float),
        _p(gw, ctypes.c_float), _p(gb, ctypes.c_float), b, c, h, w, cout, kh, kw, sh, sw, ph, pw, dh, dw, dg)
    if rc:
        raise RuntimeError(f'oracle_dcn_v2_backward rc={rc}')
    return gx, goff, gm, gw, gb
