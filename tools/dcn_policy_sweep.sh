#!/bin/bash
# experiment: L2 cache-policy variants of the DCN gather on the largest layer (time per launch)
for pol in 0 1 2 4 3 5 7; do echo "policy $pol: $(C2M_DCN_POLICY=$pol LAYER=large N=4 python tools/dcn_layers.py 2>&1 | tail -1)"; done
for mb in 64 100; do echo "policy 7 persist ${mb}MB: $(C2M_L2_PERSIST_MB=$mb LAYER=large N=4 python tools/dcn_layers.py 2>&1 | tail -1)"; done
echo "policy 1 persist 100MB: $(C2M_DCN_POLICY=1 C2M_L2_PERSIST_MB=100 LAYER=large N=4 python tools/dcn_layers.py 2>&1 | tail -1)"
