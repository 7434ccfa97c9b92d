"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, time, share.
usage: python profiles/summarize_launches.py launches.csv > summary.txt"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H, data = rows[hdr], rows[hdr + 1:]
ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
agg = {}
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    a = agg.setdefault(r[ki].split('(')[0][:90], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print('launches %d, total device time %.1f us (cold-cache, serialised under ncu: compare SHARES)' % (
    sum(a[0] for a in agg.values()), tot))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-92s n=%4d %12.1f us %6.2f%%' % (k, a[0], a[1], 100 * a[1] / tot))
