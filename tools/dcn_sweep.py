import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'c2-matching_b200')]
import torch
from c2m_b200 import ops
dev = torch.device('cuda:0')
out = []
for (B, C, H, W) in ((1, 64, 32, 32), (1, 64, 64, 64), (1, 64, 160, 160), (1, 64, 320, 320), (1, 16, 160, 160), (1, 64, 160, 16), (1, 64, 16 * 37, 8)):
    dg = 8 if C >= 64 else 2
    x = torch.randn(B, C, H, W, device=dev); w = torch.randn(C, C, 3, 3, device=dev) * 0.05; b = torch.randn(C, device=dev)
    om = torch.cat([torch.randn(B, 2 * dg * 9, H, W, device=dev) * 0.5, torch.randn(B, dg * 9, H, W, device=dev)], 1)
    pre = torch.zeros(B, 9, H, W, 2, device=dev)
    xp = ops.psa_from_f32(x)
    ops.profile_enable(True); ops.profile_collect('dcn')
    for i in range(6):
        ops.dcn_v2_fused_tc(xp, om, w, b, dg, pre_offset=pre, psa_out=True, out_f32=False)
        if i == 2: torch.cuda.synchronize(); ops.profile_collect('dcn')
    torch.cuda.synchronize()
    r = ops.profile_collect('dcn'); ops.profile_enable(False)
    tiles = B * ((H + 15) // 16) * ((W + 7) // 8)
    out.append(f'C{C} {H}x{W} ({tiles} tiles): {r["ms"] / r["launches"] * 1e3:.0f} us')
print(' | '.join(out))
