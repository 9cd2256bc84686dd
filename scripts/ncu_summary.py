"""Summarise an ncu --csv launch list: per-kernel average time, DRAM bytes, active warps."""
import csv, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); mi = hdr.index('Metric Name')
agg = OrderedDict()
for r in rows[1:]:
    if 'erd::' not in r[ki]:
        continue
    agg.setdefault(r[ki][:44], {}).setdefault(r[mi], []).append(float(r[vi].replace(',', '')))
tot = 0
avg = lambda v: sum(v) / len(v)
for k, v in agg.items():
    t = avg(v['gpu__time_duration.sum']) / 1e3
    tot += t
    rd = avg(v.get('dram__bytes_read.sum', [0])) / 1e6; wr = avg(v.get('dram__bytes_write.sum', [0])) / 1e6
    wa = avg(v.get('sm__warps_active.avg.pct_of_peak_sustained_active', [0]))
    print(f'{k:46s} n={len(v["gpu__time_duration.sum"]):3d} us={t:8.2f} dram_rd_MB={rd:8.2f} dram_wr_MB={wr:8.2f} warps_active%={wa:5.1f}')
print(f'sum of kernel times (serialised, cold): {tot:.1f} us')
