"""Summarise one .ncu-rep (single kernel, --set full) into the short text kept under profiles/."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
WANT = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_active.avg']
kn = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
print(f'kernel: {kn[:100]}')
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print(f'{w:75s} {vals[i]:>16s} {units[i]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]; data = rows[2:]
st = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
agg = {}
for r in data:
    for i in st:
        try: agg[h[i]] = agg.get(h[i], 0) + int(r[i])
        except Exception: pass
tot = sum(agg.values()) or 1
print('warp stall samples: ' + ', '.join(f'{k[6:]} {v * 100 // tot}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
