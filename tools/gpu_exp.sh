#!/bin/bash
# scratch: rescore kernel time vs chunk count and prefilter (ncu on the rescore kernels of the bench step)
for cfg in "2 1" "2 0" "4 1" "10 1" "10 0" "6 1"; do set -- $cfg
  C2M_CORR_NCHUNK=$1 C2M_RESCORE_PREFILTER=$2 B=4 STEPS=2 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rescore|corr_umma" --csv --log-file gpurun_out/r3b_$1_$2.csv python tools/one_step.py > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/r3b_$1_$2.csv') if not l.startswith('=='))]
t={}
for r in rows[len(rows)//2:]: t[r['Kernel Name'].split('(')[0]]=float(r['Metric Value'].replace(',',''))/ (1e6 if r['Metric Unit'] in ('ns','nsecond') else 1e3 if r['Metric Unit'] in ('us','usecond') else 1)
print('nchunk $1 prefilter $2', {k:round(v,3) for k,v in t.items()}, 'sum', round(sum(t.values()),3))
PY
done
