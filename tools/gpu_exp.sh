#!/bin/bash
# experiment session: DCN L2-policy sweep, per-kernel breakdown of a step, 1-CTA vs CTA-pair convolution
O=gpurun_out
bash tools/dcn_policy_sweep.sh > $O/r2e_policy.log 2>&1
python tools/gpu_breakdown.py > $O/r2e_breakdown.log 2>&1
for m in 0 1; do
  for cfg in "H=640 CIN=64 COUT=64" "H=160 CIN=64 COUT=64" "H=320 CIN=128 COUT=128" "H=640 CIN=3 COUT=64" "H=640 CIN=32 COUT=3 MODE=f32add" "H=640 CIN=64 COUT=64 MODE=res"; do
    echo "2cta=$m: $(env C2M_CONV_2CTA=$m $cfg N=5 timeout 300 python tools/conv_one.py 2>&1 | tail -1)"
  done
done > $O/r2e_conv.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -k "conv or resblock or vgg or full_forward or fast_conv" 2>&1 | tail -15 > $O/r2e_t_conv.log
cat $O/r2e_policy.log $O/r2e_conv.log; tail -3 $O/r2e_t_conv.log
