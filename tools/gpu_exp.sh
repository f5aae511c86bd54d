#!/bin/bash
O=gpurun_out
for cfg in "6 6 4" "8 12 4" "12 16 4" "12 24 4"; do set -- $cfg
  python bench.py --workload config5 --loader-workers $1 --post-workers $2 --eval-batch $3 2> $O/r2m_cfg5_$1_$2_$3.err | tail -1 > $O/r2m_cfg5_$1_$2_$3.json
  python -c "import json;j=json.load(open('$O/r2m_cfg5_$1_$2_$3.json'));print('cfg5 workers $1 post $2 batch $3:', round(j['value'],1),'img/s', j['limiting_stage'], j['per_rank'])"
done
