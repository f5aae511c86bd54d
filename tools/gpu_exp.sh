#!/bin/bash
# scratch: two-product search with top-8 lists: overflow diagnostics, parity, short bench
O=gpurun_out
C2M_CORR_DEBUG=1 B=4 STEPS=2 python tools/one_step.py 2>&1 | grep "c2m corr" | tail -2
T="python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider"
timeout 1200 $T -k "corr or feature_match or search_on or full_forward or fullsize or config2" 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-micro > $O/r2w_bench.json 2> $O/r2w_bench.err
python - <<PY
import json
d=json.load(open('$O/r2w_bench.json'))
print(round(d['value'],1),'img/s', round(d['ms_per_step'],2),'ms e2e', round(d['e2e']['value'],1), {k:round(v['ms_per_step'],2) for k,v in d['roofline']['per_kernel_class'].items()}, d['clocks']['sm_mhz'])
print({k:d['parity'][k] for k in ('idx_flips','max_gap64_of_flips','sr_max_rel_err','psnr_delta_db')})
PY
