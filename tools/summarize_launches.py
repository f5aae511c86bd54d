"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list of tools/one_step.py into the
per-kernel table kept under profiles/: one steady-state step = the launches from the last
corr_umma launch back to the one before it."""
import collections
import csv
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if not ln.startswith('==')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') == 'gpu__time_duration.sum':
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        ms = v * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}[unit]
        rows.append((r['Kernel Name'], ms))
corr = [i for i, (k, _) in enumerate(rows) if 'corr_umma_kernel' in k]
lo, hi = corr[-2], corr[-1]
step = rows[lo:hi]
acc = collections.OrderedDict()
for k, ms in step:
    k = k.split('(')[0][:80]
    a = acc.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in acc.values())
print(f'# launches in step: {len(step)}   summed device time: {tot:.2f} ms\n')
print('| kernel | launches | ms | share |\n|---|---:|---:|---:|')
for k, a in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f'| `{k}` | {a[0]} | {a[1]:.3f} | {a[1] / tot * 100:.1f}% |')
