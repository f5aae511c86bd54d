#!/bin/bash
# scratch: N1 validation variants at 1 GPU
O=gpurun_out; TAG=$1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -k "validation_gpu_metrics or cli_runs" 2>&1 | tail -60 > $O/${TAG}_t_val.log
tail -3 $O/${TAG}_t_val.log
for cfg in "6 cuda" "10 cuda" "16 cuda" "10 cpu"; do
  set -- $cfg
  python bench.py --workload config5 --loader-workers $1 --metrics-device $2 > $O/${TAG}_cfg5_$1_$2.json 2> $O/${TAG}_cfg5_$1_$2.err
  python - <<PY
import json
d=json.load(open('$O/${TAG}_cfg5_$1_$2.json'))
print('$cfg', round(d['value'],1), d['per_rank'], d['limiting_stage'])
PY
done
