"""BASELINE.json configs 3 and 4: correlation-only and DCNv2-only microbenchmarks (CUDA events,
L2 flushed between iterations), with torchvision's CUDA deform_conv2d as the stand-in comparator for
the un-buildable reference `_ext`.  Prints one JSON line per measurement."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'c2-matching_b200'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
import torch
import torchvision
import seeding
import c2m_b200 as c2m
from c2m_b200 import ops

dev = torch.device('cuda:0')
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)
peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


# ---- config 3: 256 ch, input grid h vs Ref grid hr, argmax index_map
for (h, hr) in ((40, 125), (80, 125), (125, 125), (160, 160)):
    fin = seeding.unit_features(1, 256, h, h)[None].to(dev)
    fref = seeding.unit_features(2, 256, hr, hr)[None].to(dev)
    ops.profile_enable(True); ops.profile_collect('corr_search')
    ms = timeit(lambda: c2m.corr_argmax(fin, fref, norm_input=True))
    pr = ops.profile_collect('corr_search'); ops.profile_enable(False)
    nq, nr = (h - 2) ** 2, (hr - 2) ** 2
    flops = 2.0 * 256 * 9 * nq * nr
    byts = 4.0 * 256 * (h * h + hr * hr) + 12.0 * nq
    ks = pr['ms'] / pr['launches']
    print(json.dumps({'config': 3, 'what': f'feature_match_index 256ch {h}x{h} vs {hr}x{hr}', 'ms_call': ms, 'ms_search_kernel': ks,
                      'algorithmic_tflops_kernel': flops / ks / 1e9, 'hbm_gbps_at_algorithmic_bytes_kernel': byts / ks / 1e6,
                      'frac_of_bf16_sustained': flops / ks / 1e9 / peaks.get('bf16_tflops_sustained', 1430.2),
                      'algorithmic_mb': byts / 1e6}), flush=True)

# ---- config 4: DCNv2 3x3, 64->64, 160x160, dg=8, offsets from a random index map (+ N(0,0.5) residual)
C, H, dg = 64, 160, 8
x = torch.randn(1, C, H, H, device=dev)
w = (torch.randn(C, C, 3, 3, device=dev) * 0.05)
b = torch.randn(C, device=dev)
idx = torch.randint(0, 158 * 158, (1, 158, 158), device=dev)
om = torch.cat([torch.randn(1, 2 * dg * 9, H, H, device=dev) * 0.5, torch.randn(1, dg * 9, H, H, device=dev)], 1)
pre = c2m.offset_pyramid(idx, 1)
n = dg * 9
off = om[:, :2 * n].clone(); pr_ = pre.repeat(1, dg, 1, 1, 1); off[:, 0::2] += pr_[..., 1]; off[:, 1::2] += pr_[..., 0]
mask = torch.sigmoid(om[:, 2 * n:])
xp = ops.psa_from_f32(x)
import _ext
rows = {
    'c2m dcn_umma (fused prologue, PSA in/out)': lambda: ops.dcn_v2_fused_tc(xp, om, w, b, dg, idx=idx, pre_scale=1, psa_out=True, out_f32=False),
    'c2m _ext.dcn_v2_forward (FFMA, reference boundary)': lambda: _ext.dcn_v2_forward(x, w, b, off, mask, 3, 3, 1, 1, 1, 1, 1, 1, dg),
    'torchvision.ops.deform_conv2d CUDA (comparator)': lambda: torchvision.ops.deform_conv2d(x, off, w, b, padding=1, mask=mask),
}
flops = 2.0 * C * C * 9 * H * H
byts = 4.0 * (C * H * H + 3 * dg * 9 * H * H + C * H * H) + 4 * (C * C * 9 + C)
from torch.profiler import ProfilerActivity, profile


def kernel_ms(fn, n=10):
    """Sum of device kernel time per call (CUPTI): at this size (~0.1 ms of GPU work) an event-bracketed
    call mostly measures host launch overhead, which differs between a Python/ctypes stub and a C++ op."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            flush.fill_(1)
            fn()
        torch.cuda.synchronize()
    tot = 0.0
    for ev in prof.key_averages():
        if 'fill' in ev.key.lower() or 'FillFunctor' in ev.key:
            continue
        tot += getattr(ev, 'self_device_time_total', getattr(ev, 'self_cuda_time_total', 0.0))
    return tot / n / 1e3


for name, fn in rows.items():
    ms_call = timeit(fn)
    ms_k = kernel_ms(fn)
    print(json.dumps({'config': 4, 'what': name, 'ms_call_events': ms_call, 'ms_kernels': ms_k,
                      'algorithmic_tflops_kernels': flops / ms_k / 1e9, 'hbm_gbps_at_algorithmic_bytes_kernels': byts / ms_k / 1e6,
                      'algorithmic_mb': byts / 1e6}), flush=True)
