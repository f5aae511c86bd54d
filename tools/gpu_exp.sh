#!/bin/bash
O=gpurun_out
for m in 0 1; do
  for cfg in "H=640 CIN=64 COUT=64" "H=160 CIN=64 COUT=64" "H=320 CIN=128 COUT=128" "H=640 CIN=64 COUT=64 MODE=res" "H=640 CIN=3 COUT=64" "H=640 CIN=32 COUT=3 MODE=f32add"; do
    echo "2cta=$m: $(env C2M_CONV_2CTA=$m $cfg N=5 timeout 300 python tools/conv_one.py 2>&1 | tail -1)"
  done
done > $O/r2h_conv.log 2>&1
LAYER=all N=3 python tools/dcn_layers.py >> $O/r2h_conv.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -k "not corr and not feature_match and not search_on" 2>&1 | tail -5 > $O/r2h_t.log
C2M_CONV_2CTA=0 python tools/gpu_breakdown.py > $O/r2h_breakdown.log 2>&1
cat $O/r2h_conv.log; tail -3 $O/r2h_t.log; grep -E "ms/step|c2m::" $O/r2h_breakdown.log | cut -c1-60,150-250 | head -8
