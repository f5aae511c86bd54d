#!/bin/bash
# scratch: CUDA-graph replay: parity test + bench with / without
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -k "cuda_graph or deterministic" 2>&1 | tail -15
for G in 1 0; do
python bench.py --steps 10 --warmup 3 --no-micro --no-cpu-baseline --cuda-graph $G > $O/r2x_bench_g$G.json 2> $O/r2x_bench_g$G.err; tail -2 $O/r2x_bench_g$G.err
python - <<PY
import json
d=json.load(open('$O/r2x_bench_g$G.json'))
print('graph=$G', round(d['value'],1),'img/s', round(d['ms_per_step'],2),'ms e2e', round(d['e2e']['value'],1), d['gpu_launches'], d['clocks']['sm_mhz'])
PY
done
