#!/bin/bash
O=gpurun_out
for cfg in "H=640 CIN=64 COUT=64" "H=640 CIN=64 COUT=64 MODE=res" "H=160 CIN=64 COUT=64 MODE=res" "H=320 CIN=64 COUT=64 MODE=res"; do
  echo "$(env $cfg N=5 timeout 300 python tools/conv_one.py 2>&1 | tail -1)"
done > $O/r2j_conv.log 2>&1
for pol in 5 13; do echo "policy $pol:"; C2M_DCN_POLICY=$pol LAYER=all N=3 python tools/dcn_layers.py 2>&1 | tail -3; done >> $O/r2j_conv.log 2>&1
cat $O/r2j_conv.log
bash tools/gpu_check.sh r2j tests benchfull
LAYER=small N=1 ncu --set full --clock-control none --import-source on -k regex:dcn_umma -s 2 -c 1 -f -o $O/r2j_dcn_small python tools/dcn_layers.py > $O/r2j_ncu_dcn.log 2>&1
