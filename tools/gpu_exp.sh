#!/bin/bash
# scratch: A/B of the MMA issue path on one box (dbg 64 = per-MMA elect (old), else one elected thread per tile)
for rep in 1 2; do for MODE in psa res; do for D in 0 64 4 68 12 76; do
  echo -n "MODE=$MODE DBG=$D  "; MODE=$MODE C2M_CONV_DBG=$D N=9 python tools/conv_one.py 2>&1 | tail -1
done; done; done
