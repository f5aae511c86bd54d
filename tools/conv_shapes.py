"""Per-shape device time of the conv3x3 launches of one bench step (library CUDA-event hook, one
sync per launch — diagnostic only)."""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'c2-matching_b200'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from c2m_b200 import ops  # noqa: E402
from c2m_b200.pipeline import RestorationPipeline, synthetic_pair  # noqa: E402

B = int(os.environ.get('B', '4'))
dev = torch.device('cuda:0')
pipe = RestorationPipeline(dev).load_state_dicts(*bench.seeded_weights()).place()
x = [t.to(dev) for t in synthetic_pair(1234, B, 160, 500)]
for _ in range(2):
    pipe.forward(*x)
torch.cuda.synchronize()
acc = collections.OrderedDict()
orig = ops.conv3x3_psa


def wrapped(xp, weight, bias, **kw):
    ops.profile_collect('conv3x3')
    r = orig(xp, weight, bias, **kw)
    torch.cuda.synchronize()
    rec = ops.profile_collect('conv3x3')
    cin = weight.shape[1]
    key = (xp.B, cin, weight.shape[0], xp.H, xp.W, 'ps' if kw.get('pixel_shuffle') else '', 'f32' if kw.get('out_f32') else '',
           'res' if kw.get('residual') is not None else '')
    a = acc.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += rec['ms']
    a[2] += rec['flops']
    return r


ops.conv3x3_psa = wrapped
ops.profile_enable(True)
pipe.forward(*x)
torch.cuda.synchronize()
ops.profile_enable(False)
tot = sum(a[1] for a in acc.values())
print(f'conv3x3 total {tot:.2f} ms over {sum(a[0] for a in acc.values())} launches')
for k, a in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    tiles = k[0] * ((k[3] + 15) // 16) * ((k[4] + 7) // 8)
    print(f'{str(k):58s} n={a[0]:3d} ms={a[1]:7.3f} ({a[1] / a[0] * 1e3:7.1f} us each) {a[2] / a[1] / 1e9:7.1f} TF/s alg  tiles={tiles}')
