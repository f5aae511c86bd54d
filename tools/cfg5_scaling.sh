#!/bin/bash
# config-5 (sharded evaluation) strong-scaling curve on ONE box: N = 1, 2, 4, 8.   usage: bash tools/cfg5_scaling.sh <tag> [Ns...]
TAG=$1; shift; O=gpurun_out
for N in ${@:-1 2 4 8}; do
  if [ $N = 1 ]; then
    python bench.py --workload config5 > $O/${TAG}_cfg5_g$N.json 2> $O/${TAG}_cfg5_g$N.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) \
      bench.py --gpus $N --workload config5 2> $O/${TAG}_cfg5_g$N.err | tail -1 > $O/${TAG}_cfg5_g$N.json
  fi
  python - <<PY
import json
d=json.load(open('$O/${TAG}_cfg5_g$N.json'))
print($N, round(d['value'],1), 'img/s', round(d['ms_per_step']), 'ms', d['limiting_stage'], d['per_rank'][:2])
PY
done
