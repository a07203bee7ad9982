#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of counters the roofline notes quote."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_fma.sum', 'smsp__inst_executed_pipe_alu.sum',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    for r in rows[2:]:
        print('---', r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
        for k in KEYS:
            if k in hdr:
                print('  %-70s %-14s %s' % (k, rows[1][hdr.index(k)], r[hdr.index(k)]))
        for i, h in enumerate(hdr):
            if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct'):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 3:
                    print('  STALL %-40s %.1f' % (h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''), v))


if __name__ == '__main__':
    main()
