"""Experiment: is the forward step launch-bound anywhere?  Eager (CPU issue time, device time) vs a CUDA-graph replay."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'c2-matching_b200'), os.path.join(ROOT, 'tests', 'golden')):
    sys.path.insert(0, p)
import torch  # noqa: E402

import bench  # noqa: E402
from c2m_b200.pipeline import RestorationPipeline, synthetic_pair  # noqa: E402

B = int(os.environ.get('B', '4'))
dev = torch.device('cuda:0')
pipe = RestorationPipeline(dev).load_state_dicts(*bench.seeded_weights()).place()
x = [t.to(dev) for t in synthetic_pair(1234, B, 160, 500)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())


def timed(fn, n=8):
    cpu, gpu = [], []
    for _ in range(n):
        flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t = time.perf_counter()
        a.record()
        fn()
        b.record()
        cpu.append(time.perf_counter() - t)
        torch.cuda.synchronize()
        gpu.append(a.elapsed_time(b))
    cpu.sort(), gpu.sort()
    return cpu[len(cpu) // 2] * 1e3, gpu[len(gpu) // 2]


with torch.cuda.stream(s):
    for _ in range(3):
        ref = pipe.forward(*x)
    torch.cuda.synchronize()
    print('eager  cpu-issue %.2f ms   device %.2f ms' % timed(lambda: pipe.forward(*x)))
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, stream=s):
            out = pipe.forward(*x)
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        print('graph  bit-identical to eager:', bool(torch.equal(out, ref)))
        print('graph  cpu-issue %.2f ms   device %.2f ms' % timed(g.replay))
    except Exception as e:
        print('capture failed:', repr(e)[:600])
