#!/usr/bin/env python
"""Compact per-kernel summary of an `ncu --set full` report:  python profiles/summarize_ncu.py rep.ncu-rep > out.txt"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed',
        'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']


def main():
    rep = sys.argv[1]
    if rep.endswith('.csv'):      # raw page already exported on the GPU box (ncu -i rep --page raw --csv)
        out = open(rep).read()
    else:
        out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u = rows[0], rows[1]
    idx = {k: i for i, k in enumerate(h)}
    print('# source: %s (ncu --set full --clock-control none)' % rep.split('/')[-1])
    for r in rows[2:]:
        print('kernel: %s' % r[idx['Kernel Name']])
        for k in KEYS:
            if k in idx:
                print('  %-86s %-14s %s' % (k, u[idx[k]], r[idx[k]]))
        print()


if __name__ == '__main__':
    main()
