#!/usr/bin/env python
"""DRAM traffic and tensor-pipe activity of the tcgen05 Conv1D launches of ONE training step, from
    ncu --set full --clock-control none -k regex:conv_tc -s 63 -c 21 -o rep python bench.py --steps 2 --warmup 3
usage: python profiles/summarize_step_traffic.py rep.ncu-rep profiles/rNN_conv_tc_step_traffic.json"""
import csv
import io
import json
import subprocess
import sys


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    if rep.endswith('.csv'):      # already exported on the GPU box: ncu -i rep --page raw --csv > file.csv
        out = open(rep).read()
    else:
        out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u = rows[0], rows[1]
    idx = {k: i for i, k in enumerate(h)}
    mult = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, '%': 1}

    def val(r, k):
        return float(r[idx[k]].replace(',', '')) * mult.get(u[idx[k]], 1)
    per = []
    for r in rows[2:]:
        per.append({'kernel': r[idx['Kernel Name']].split('(')[0],
                    'dram_read_bytes': val(r, 'dram__bytes_read.sum'), 'dram_write_bytes': val(r, 'dram__bytes_write.sum'),
                    'time_s': val(r, 'gpu__time_duration.sum'),
                    'tensor_pipe_active_pct': val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed')})
    tot = sum(p['dram_read_bytes'] + p['dram_write_bytes'] for p in per)
    tt = sum(p['time_s'] for p in per)
    what = sys.argv[3] if len(sys.argv) > 3 else ('the tcgen05 Conv1D launches (fwd, dgrad, wgrad of the 7 tensor-core layers) of '
                                                  'one signal_pe training step, batch 512, n_pix 2048')
    src = sys.argv[4] if len(sys.argv) > 4 else 'ncu --set full --clock-control none -k regex:conv_tc -s 63 -c 21 python bench.py --steps 2 --warmup 3'
    res = {'source': '%s (%s)' % (src, rep.split('/')[-1]),
           'what': what,
           'launches': len(per), 'dram_bytes_per_step': tot, 'dram_bytes_per_launch': tot / max(len(per), 1),
           'time_weighted_tensor_pipe_active_pct': sum(p['time_s'] * p['tensor_pipe_active_pct'] for p in per) / tt,
           'per_launch': per}
    with open(dst, 'w') as f:
        json.dump(res, f, indent=1)
    for p in per:
        print('%-44s %8.1f MB rd %8.1f MB wr %8.1f us  tensor pipe %.1f%%' % (
            p['kernel'], p['dram_read_bytes'] / 1e6, p['dram_write_bytes'] / 1e6, p['time_s'] * 1e6, p['tensor_pipe_active_pct']))
    print('%d launches, %.2f GB DRAM per step, time-weighted tensor pipe active %.1f%%' % (
        len(per), tot / 1e9, res['time_weighted_tensor_pipe_active_pct']))


if __name__ == '__main__':
    main()
