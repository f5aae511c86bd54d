#!/bin/bash
# scratch: conv timing after the uniform-datapath MMA issue + parity subset + short bench
for MODE in psa res; do MODE=$MODE N=9 python tools/conv_one.py 2>&1 | tail -1; done
H=320 MODE=res N=9 python tools/conv_one.py | tail -1
H=160 MODE=res N=9 python tools/conv_one.py | tail -1
H=320 CIN=128 COUT=128 N=9 python tools/conv_one.py | tail -1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -k "conv3x3 or fast_conv or full_forward" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-micro 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['value'],1),'img/s', round(d['ms_per_step'],2),'ms', {k:round(v['ms_per_step'],2) for k,v in d['roofline']['per_kernel_class'].items()}, d['clocks'])"
