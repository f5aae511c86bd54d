#!/bin/bash
# ncu launch list (time + DRAM bytes per launch) of two bench steps -> gpurun_out/<tag>_launches.csv
TAG=${1:-r2}
B=4 STEPS=3 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/${TAG}_launches.csv python tools/one_step.py > gpurun_out/${TAG}_launches.log 2>&1
tail -2 gpurun_out/${TAG}_launches.log; wc -l gpurun_out/${TAG}_launches.csv
