#!/usr/bin/env python
"""Condense ncu outputs into the small text summaries committed under profiles/.

  python tools/ncu_summary.py full  gpurun_out/prof_X.ncu-rep            > profiles/rNN_X_full.txt
  python tools/ncu_summary.py launches gpurun_out/launches_X.csv         > profiles/rNN_X_launches.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
    'launch__shared_mem_per_block_dynamic', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum',
    'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'smsp__thread_inst_executed_per_inst_executed.ratio', 'local_load', 'local_store',
]


def full(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print('# ncu --set full --clock-control none  (%s), %d captured launches' % (rep, len(data)))
    ki = hdr.index('Kernel Name')
    for r in data:
        print('kernel: %s   grid %s block %s' % (r[ki], r[hdr.index('Grid Size')], r[hdr.index('Block Size')]))
    for i, h in enumerate(hdr):
        if any(h == k or (k in h and k in ('local_load', 'local_store')) for k in KEYS) or \
                ('issue_stalled' in h and h.endswith('per_issue_active.ratio')):
            print('%-90s %-12s %s' % (h, units[i], '  '.join(r[i] for r in data)))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    tot = collections.OrderedDict()
    for r in rows[1:]:
        name = r[kn]
        short = name if len(name) < 110 else name[:107] + '...'
        t = tot.setdefault(short, [0, 0.0])
        t[0] += 1
        t[1] += float(r[mv].replace(',', ''))
    total = sum(v[1] for v in tot.values())
    print('# ncu --metrics gpu__time_duration.sum --clock-control none  (%s): %d launches, %.3f ms total' % (path, len(rows) - 1, total / 1e6))
    print('# share   launches   total_us   avg_us   kernel')
    for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print('%6.2f%% %8d %10.1f %9.1f   %s' % (100 * ns / total, n, ns / 1e3, ns / 1e3 / n, k))


if __name__ == '__main__':
    {'full': full, 'launches': launches}[sys.argv[1]](sys.argv[2])
