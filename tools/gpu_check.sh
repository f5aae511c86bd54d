#!/bin/bash
# One GPU-box session: parity tests in three processes (a trapped kernel then only takes its own group down),
# per-kernel timings, the bench, optional ncu captures.  usage: bash tools/gpu_check.sh <tag> [tests] [bench|benchfull] [sanitize] [cfg5] [ncu_corr] [ncu_dcn] [ncu_conv]
TAG=$1; shift
O=gpurun_out
T="python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider"
NC="ncu --set full --clock-control none --import-source on"
for what in "$@"; do
  case $what in
    tests)
      timeout 1200 $T -k "corr or feature_match or search_on" 2>&1 | tail -25 > $O/${TAG}_t_corr.log
      timeout 1200 $T -k "dcn" 2>&1 | tail -25 > $O/${TAG}_t_dcn.log
      timeout 1500 $T -k "not corr and not feature_match and not search_on and not dcn" 2>&1 | tail -25 > $O/${TAG}_t_rest.log
      tail -2 $O/${TAG}_t_*.log ;;
    kern)
      python tools/dcn_layers.py > $O/${TAG}_dcn_layers.log 2>&1
      B=4 python tools/corr_one.py > $O/${TAG}_corr_one.log 2>&1
      cat $O/${TAG}_dcn_layers.log $O/${TAG}_corr_one.log ;;
    bench) python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-micro > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err ;;
    benchfull) python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err ;;
    sanitize)
      timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > $O/${TAG}_memcheck.log 2>&1; echo memcheck rc=$?
      timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_small.py > $O/${TAG}_synccheck.log 2>&1; echo synccheck rc=$?
      tail -3 $O/${TAG}_memcheck.log $O/${TAG}_synccheck.log ;;
    cfg5) python bench.py --workload config5 > $O/${TAG}_cfg5.json 2> $O/${TAG}_cfg5.err; tail -2 $O/${TAG}_cfg5.err ;;
    ncu_corr) B=4 N=2 $NC -k regex:corr_umma -s 1 -c 1 -f -o $O/${TAG}_corr python tools/corr_one.py > $O/${TAG}_ncu_corr.log 2>&1 ;;
    ncu_dcn) LAYER=large N=1 $NC -k regex:dcn_umma -s 2 -c 1 -f -o $O/${TAG}_dcn_large python tools/dcn_layers.py > $O/${TAG}_ncu_dcn.log 2>&1 ;;
    ncu_conv) N=3 $NC -k regex:conv3x3_umma -s 2 -c 1 -f -o $O/${TAG}_conv64 python tools/conv_one.py > $O/${TAG}_ncu_conv.log 2>&1 ;;
  esac
done
