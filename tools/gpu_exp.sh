#!/bin/bash
# scratch: chunk-count model with the re-scan term: diagnostics, corr parity, bench
O=gpurun_out
C2M_CORR_DEBUG=1 B=4 STEPS=2 python tools/one_step.py 2>&1 | grep "c2m corr" | tail -1
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -k "corr or feature_match or search_on or fullsize or config2" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2z_bench.json 2> $O/r2z_bench.err
python - <<PY
import json
d=json.load(open('$O/r2z_bench.json'))
print(round(d['value'],1),'img/s', round(d['ms_per_step'],2),'ms e2e', round(d['e2e']['value'],1), {k:round(v['ms_per_step'],2) for k,v in d['roofline']['per_kernel_class'].items()}, d['clocks']['sm_mhz'])
for m in d['micro']['rows'] if 'rows' in d['micro'] else d['micro'].get('results', []):
    if 'ms_search_kernel' in m: print(m['what'], round(m['ms_call'],3), round(m['ms_search_kernel'],4))
PY
