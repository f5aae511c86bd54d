"""BASELINE.json configs 3 and 4: correlation-only and DCNv2-only microbenchmarks (CUDA events / CUPTI kernel
time, L2 flushed between iterations), with torchvision's CUDA deform_conv2d as the stand-in comparator for the
un-buildable reference `_ext`.  `bench.py` imports run_config3 / run_config4 for its `micro`
block (so the driver runs them); as a script it prints one JSON line per measurement."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'c2-matching_b200'), os.path.join(ROOT, 'tests', 'golden')):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch


def _timeit(fn, flush, n=20, warm=3):
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


def _kernel_ms(fn, flush, n=10):
    """Sum of device kernel time per call (CUPTI): at this size (~0.1 ms of GPU work) an event-bracketed
    call mostly measures host launch overhead, which differs between a Python/ctypes stub and a C++ op."""
    from torch.profiler import ProfilerActivity, profile
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


def run_config3(dev, flush, peak_tflops, shapes=((40, 125), (80, 125), (125, 125), (160, 160)), n=20):
    """256 ch, input grid h vs Ref grid hr, argmax index_map: time of the candidate-search kernel alone (library
    CUDA-event hook) and of the whole feature_match_index call (prep + search + exact rescoring)."""
    import seeding
    import c2m_b200 as c2m
    from c2m_b200 import ops
    rows = []
    for (h, hr) in shapes:
        fin = seeding.unit_features(1, 256, h, h)[None].to(dev)
        fref = seeding.unit_features(2, 256, hr, hr)[None].to(dev)
        ops.profile_enable(True); ops.profile_collect('corr_search')
        ms = _timeit(lambda: c2m.corr_argmax(fin, fref, norm_input=True), flush, n=n)
        pr = ops.profile_collect('corr_search'); ops.profile_enable(False)
        nq, nr = (h - 2) ** 2, (hr - 2) ** 2
        flops = 2.0 * 256 * 9 * nq * nr
        byts = 4.0 * 256 * (h * h + hr * hr) + 12.0 * nq
        ks = pr['ms'] / pr['launches']
        rows.append({'config': 3, 'what': f'feature_match_index 256ch {h}x{h} vs {hr}x{hr}', 'ms_call': ms,
                     'ms_search_kernel': ks, 'algorithmic_tflops_kernel': flops / ks / 1e9,
                     'hbm_gbps_at_algorithmic_bytes_kernel': byts / ks / 1e6,
                     'frac_of_bf16_peak': flops / ks / 1e9 / peak_tflops, 'algorithmic_mb': byts / 1e6,
                     # nine-tap fp32-equivalent flops vs what the kernel issues: 3 row taps x 2 fp16 products x (16/14)^2 halo
                     'issued_over_algorithmic_mma': round(2.0 / 3.0 * (16.0 / 14.0) ** 2, 3)})
    return rows


def run_config4(dev, flush, hbm_gbs, n=20):
    """DCNv2 3x3, 64->64, 160x160, dg=8, offsets from a random index map (+ N(0,0.5) residual): the fused kernel
    (prologue fused, operand layouts in/out), the literal `_ext.dcn_v2_forward` boundary (fp32 NCHW in/out,
    conversions included) and torchvision's CUDA op on the same tensors."""
    import torchvision
    import _ext
    import c2m_b200 as c2m
    from c2m_b200 import ops
    C, H, dg = 64, 160, 8
    g = torch.Generator(device='cpu').manual_seed(4)
    x = torch.randn(1, C, H, H, generator=g).to(dev)
    w = (torch.randn(C, C, 3, 3, generator=g) * 0.05).to(dev)
    b = torch.randn(C, generator=g).to(dev)
    idx = torch.randint(0, 158 * 158, (1, 158, 158), generator=g).to(dev)
    om = torch.cat([torch.randn(1, 2 * dg * 9, H, H, generator=g) * 0.5, torch.randn(1, dg * 9, H, H, generator=g)], 1).to(dev)
    pre = c2m.offset_pyramid(idx, 1)
    nn_ = dg * 9
    off = om[:, :2 * nn_].clone(); pr_ = pre.repeat(1, dg, 1, 1, 1); off[:, 0::2] += pr_[..., 1]; off[:, 1::2] += pr_[..., 0]
    mask = torch.sigmoid(om[:, 2 * nn_:])
    xp = ops.psa_from_f32(x)
    c8 = (27 * dg + 7) // 8
    padded = torch.zeros(1, c8 * 8, H, H, device=dev); padded[:, :27 * dg] = om
    om_oct = ops.OctF32(padded.view(1, c8, 8, H, H).permute(0, 1, 3, 4, 2).contiguous(), 27 * dg)
    fns = {
        'c2m dcn_umma (fused prologue, operand layouts in/out)':
            lambda: ops.dcn_v2_fused_tc(xp, om_oct, w, b, dg, idx=idx, pre_scale=1, psa_out=True, out_f32=False),
        'c2m _ext.dcn_v2_forward (reference boundary: fp32 NCHW in/out, conversions included)':
            lambda: _ext.dcn_v2_forward(x, w, b, off, mask, 3, 3, 1, 1, 1, 1, 1, 1, dg),
        'torchvision.ops.deform_conv2d CUDA (comparator)':
            lambda: torchvision.ops.deform_conv2d(x, off, w, b, padding=1, mask=mask),
    }
    ref = fns['torchvision.ops.deform_conv2d CUDA (comparator)']()
    got = fns['c2m _ext.dcn_v2_forward (reference boundary: fp32 NCHW in/out, conversions included)']()
    agree = float((got - ref).abs().max() / ref.abs().max())
    flops = 2.0 * C * C * 9 * H * H
    byts = 4.0 * (C * H * H + 3 * dg * 9 * H * H + C * H * H) + 4 * (C * C * 9 + C)
    rows = []
    for name, fn in fns.items():
        ms_call = _timeit(fn, flush, n=n)
        ms_k = _kernel_ms(fn, flush)
        rows.append({'config': 4, 'what': name, 'ms_call_events': ms_call, 'ms_kernels': ms_k,
                     'algorithmic_tflops_kernels': flops / ms_k / 1e9, 'hbm_gbps_at_algorithmic_bytes_kernels': byts / ms_k / 1e6,
                     'frac_of_hbm_peak': byts / ms_k / 1e6 / hbm_gbs, 'algorithmic_mb': byts / 1e6})
    rows.append({'config': 4, 'what': '_ext vs torchvision max rel diff', 'value': agree})
    return rows


if __name__ == '__main__':
    dev = torch.device('cuda:0')
    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    peaks = json.load(open(pk)) if os.path.exists(pk) else {}
    for r in run_config3(dev, flush, peaks.get('bf16_tflops', 1590.0)) + run_config4(dev, flush, peaks.get('hbm_gbs', 6650.0)):
        print(json.dumps(r), flush=True)
