#!/bin/bash
O=gpurun_out
for il in 0 1; do echo "interleave=$il:"; C2M_DCN_INTERLEAVE=$il LAYER=all N=3 python tools/dcn_layers.py 2>&1 | tail -3; done > $O/r2o_dcn.log 2>&1
python tools/micro4.py 2>/dev/null | cut -c1-200 >> $O/r2o_dcn.log
cat $O/r2o_dcn.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -k "dcn" 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-micro > $O/r2o_bench.json 2> $O/r2o_bench.err
python -c "
import json;j=json.loads(open('$O/r2o_bench.json').read().strip().splitlines()[-1]);print(j['value'], j['e2e']['value']);[print(k,round(v['ms_per_step'],2)) for k,v in j['roofline']['per_kernel_class'].items()]"
