"""One 3x3 convolution shape in isolation (for ncu captures / quick timing): env H (640), CIN (64), COUT (64), B (4),
MODE in {psa (default), res (PSA residual), f32add (fp32 NCHW out + added tensor, as the last layer)}, N launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'c2-matching_b200')]
import torch
from c2m_b200 import ops
H, CIN, COUT, B = (int(os.environ.get(k, d)) for k, d in (('H', '640'), ('CIN', '64'), ('COUT', '64'), ('B', '4')))
MODE, N = os.environ.get('MODE', 'psa'), int(os.environ.get('N', '3'))
dev = torch.device('cuda:0')
x = torch.randn(B, CIN, H, H, device=dev)
w = torch.randn(COUT, CIN, 3, 3, device=dev) * 0.05
b = torch.randn(COUT, device=dev)
xp = ops.psa_from_f32(x)
kw = {}
if MODE == 'res':
    kw = dict(residual=ops.psa_from_f32(torch.randn(B, COUT, H, H, device=dev)), out=ops.PSA.empty(B, COUT, H, H, dev))
elif MODE == 'f32add':
    kw = dict(psa_out=False, out_f32=True, add_f32=torch.randn(B, COUT, H, H, device=dev))
else:
    kw = dict(out=ops.PSA.empty(B, COUT, H, H, dev))
flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(N)]
for e0, e1 in ev:
    flush.fill_(1)
    e0.record()
    ops.conv3x3_psa(xp, w, b, act='relu' if MODE != 'f32add' else None, **kw)
    e1.record()
torch.cuda.synchronize()
ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
print(f'conv {CIN}->{COUT} @{H}x{H} B={B} {MODE}: median {ms[len(ms) // 2] * 1e3:.1f} us, min {ms[0] * 1e3:.1f} us over {N} launches')
