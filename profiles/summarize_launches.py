"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, time, share.
usage: python profiles/summarize_launches.py launches.csv [exclude-substring ...] > summary.txt
Kernels whose name contains one of the exclude substrings are listed but left out of a second set of shares (the bench
command also times gn_whiten_td_f32 alone on 8192 / 32768 series, which is not part of a training step)."""
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
excl = sys.argv[2:]
tot2 = sum(a[1] for k, a in agg.items() if not any(e in k for e in excl))
print('launches %d, total device time %.1f us (cold-cache, serialised under ncu: compare SHARES)' % (
    sum(a[0] for a in agg.values()), tot))
if excl:
    print('second share column: without %s (%.1f us left)' % (', '.join(excl), tot2))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    skip = any(e in k for e in excl)
    extra = '' if not excl else ('   (excluded)' if skip else ' %6.2f%%' % (100 * a[1] / tot2))
    print('%-92s n=%4d %12.1f us %6.2f%%%s' % (k, a[0], a[1], 100 * a[1] / tot, extra))
