#!/bin/bash
# scratch: rescore kernels after moving the score helpers out of line
for cfg in "4" "10"; do
  C2M_CORR_NCHUNK=$cfg B=4 STEPS=2 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rescore|corr_umma" --csv --log-file gpurun_out/r3f_$cfg.csv python tools/one_step.py > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/r3f_$cfg.csv') if not l.startswith('=='))]
t={}
for r in rows[len(rows)//2:]: t[r['Kernel Name'].split('(')[0]]=float(r['Metric Value'].replace(',',''))/ (1e6 if r['Metric Unit'] in ('ns','nsecond') else 1e3 if r['Metric Unit'] in ('us','usecond') else 1)
print('nchunk $cfg', {k:round(v,3) for k,v in t.items()}, 'sum', round(sum(t.values()),3))
PY
done
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider -k "corr or feature_match or search_on or fullsize or config2" 2>&1 | tail -3
