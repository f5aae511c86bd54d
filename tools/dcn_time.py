import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'c2-matching_b200')]
import torch
import c2m_b200 as c2m
from c2m_b200 import ops
dev = torch.device('cuda:0')
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)
out = []
for (B, C, H, rnd) in ((1, 64, 160, 1), (1, 64, 160, 0), (4, 64, 160, 1), (1, 64, 640, 0)):
    dg = 8
    x = torch.randn(B, C, H, H, device=dev); w = torch.randn(C, C, 3, 3, device=dev) * 0.05; b = torch.randn(C, device=dev)
    gh = 158; sc = H // 160
    if rnd:
        idx = torch.randint(0, gh * gh, (B, gh, gh), device=dev)
    else:   # coherent: near-identity flow + smooth shift
        base = torch.arange(gh * gh, device=dev).view(1, gh, gh).repeat(B, 1, 1)
        idx = (base + 3 * gh + 5).clamp_(0, gh * gh - 1)
    om = torch.cat([torch.randn(B, 144, H, H, device=dev) * 0.5, torch.randn(B, 72, H, H, device=dev)], 1)
    xp = ops.psa_from_f32(x)
    ops.profile_enable(True); ops.profile_collect('dcn')
    for i in range(8):
        if i >= 3: flush.fill_(1)
        ops.dcn_v2_fused_tc(xp, om, w, b, dg, idx=idx, pre_scale=sc, psa_out=True, out_f32=False)
        if i == 2: torch.cuda.synchronize(); ops.profile_collect('dcn')
    torch.cuda.synchronize()
    r = ops.profile_collect('dcn'); ops.profile_enable(False)
    out.append(f'B{B} C{C} {H}x{H} {"rand" if rnd else "coh"}: {r["ms"] / r["launches"]:.3f} ms')
print(' | '.join(out))
