import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'c2-matching_b200')]
import torch
from c2m_b200 import ops
dev = torch.device('cuda:0')
B, C, H, W, dg = 1, 64, 32, 32, 8
x = torch.randn(B, C, H, W, device=dev); w = torch.randn(C, C, 3, 3, device=dev) * 0.05; b = torch.randn(C, device=dev)
om = torch.cat([torch.randn(B, 2 * dg * 9, H, W, device=dev) * 0.5, torch.randn(B, dg * 9, H, W, device=dev)], 1)
pre = torch.zeros(B, 9, H, W, 2, device=dev)
xp = ops.psa_from_f32(x)
for i in range(3):
    ops.dcn_v2_fused_tc(xp, om, w, b, dg, pre_offset=pre, psa_out=True, out_f32=False)
    torch.cuda.synchronize()
    print('---', flush=True)
