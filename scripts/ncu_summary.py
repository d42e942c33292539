#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/.

  ncu_summary.py launches <launches.csv>          per-kernel time share
  ncu_summary.py full <report.ncu-rep> [kernel-id] key metrics of a --set full capture
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'launch__waves_per_multiprocessor',
    'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__inst_executed.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[h]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    d = defaultdict(list)
    for r in rows[h + 1:]:
        if len(r) > vi:
            try:
                d[r[ki].split('(')[0]].append(float(r[vi].replace(',', '')))
            except ValueError:
                pass
    tot = sum(sum(v) for v in d.values())
    print("kernel                                                        launches   total_ms  share")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print('%-62s %6d %10.3f %5.1f%%' % (k[:62], len(v), sum(v) / 1e6, 100 * sum(v) / tot))


def full(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
        print('== kernel:', name)
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print('%-78s %s %s' % (k, vals[i], units[i]))


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
