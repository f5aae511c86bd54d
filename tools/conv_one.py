"""One 64->64 3x3 conv at 640x640, B=4 (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'c2-matching_b200')]
import torch
from c2m_b200 import ops
H = int(os.environ.get('H', '640'))
dev = torch.device('cuda:0')
x = torch.randn(4, 64, H, H, device=dev)
w = torch.randn(64, 64, 3, 3, device=dev) * 0.05
b = torch.randn(64, device=dev)
xp = ops.psa_from_f32(x)
yp = ops.PSA.empty(4, 64, H, H, dev)
for _ in range(3):
    ops.conv3x3_psa(xp, w, b, act='relu', out=yp)
torch.cuda.synchronize()
