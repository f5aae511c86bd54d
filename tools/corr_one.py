"""One batched correlation search at BASELINE config-2 map size (for ncu captures / quick timing): env B (1), H (160)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'c2-matching_b200'), os.path.join(ROOT, 'tests', 'golden')]
import torch
import c2m_b200 as c2m
from c2m_b200 import ops
B, H, N = int(os.environ.get('B', '1')), int(os.environ.get('H', '160')), int(os.environ.get('N', '3'))
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(3)
fin = torch.nn.functional.normalize(torch.randn(B, 256, H, H, generator=g), dim=1).to(dev)
fref = torch.nn.functional.normalize(torch.randn(B, 256, H, H, generator=g), dim=1).to(dev)
ops.profile_enable(True)
for _ in range(N):
    idx, val = c2m.corr_argmax(fin, fref, norm_input=True)
torch.cuda.synchronize()
pr = ops.profile_collect('corr_search')
print(f'corr search B={B} {H}x{H}: {pr["ms"] / pr["launches"]:.3f} ms per launch, {pr["flops"] / pr["ms"] / 1e9:.1f} TFLOP/s algorithmic')
