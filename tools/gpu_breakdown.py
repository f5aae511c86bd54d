"""Per-kernel device-time breakdown of one bench step (torch.profiler / CUPTI), to decide what to
optimise next.  Diagnostic only — numbers under a profiler are never bench values."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'c2-matching_b200'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from c2m_b200.pipeline import RestorationPipeline, synthetic_pair  # noqa: E402

tf32 = int(os.environ.get('TF32', '0'))
cl = int(os.environ.get('CL', '0'))
B = int(os.environ.get('B', '4'))
dev = torch.device('cuda:0')
pipe = RestorationPipeline(dev, allow_tf32=bool(tf32), channels_last=bool(cl)).load_state_dicts(*bench.seeded_weights()).place()
x = [t.to(dev) for t in synthetic_pair(1234, B, 160, 500)]
for _ in range(3):
    pipe.forward(*x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    pipe.forward(*x)
e1.record()
torch.cuda.synchronize()
print(f'tf32={tf32} channels_last={cl} B={B}: {e0.elapsed_time(e1) / 3:.2f} ms/step  ({B * 3 / e0.elapsed_time(e1) * 1e3:.2f} img/s)')
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    pipe.forward(*x)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=70))
