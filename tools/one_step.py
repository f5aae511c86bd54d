"""One warm-up + one measured forward step of the bench workload (for ncu captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'c2-matching_b200'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from c2m_b200.pipeline import RestorationPipeline, synthetic_pair  # noqa: E402

B = int(os.environ.get('B', '4'))
steps = int(os.environ.get('STEPS', '2'))
dev = torch.device('cuda:0')
pipe = RestorationPipeline(dev, allow_tf32=bool(int(os.environ.get('TF32', '0')))).load_state_dicts(*bench.seeded_weights()).place()
x = [t.to(dev) for t in synthetic_pair(1234, B, 160, 500)]
for _ in range(steps):
    pipe.forward(*x)
    torch.cuda.synchronize()
print('done')
