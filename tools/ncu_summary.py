"""Prints the handful of ncu raw-page metrics we track for the trace kernel.  usage: ncu_summary.py raw.csv [kernel-row]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
vals = rows[2 + (int(sys.argv[2]) if len(sys.argv) > 2 else 0)]
want = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit', 'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed', 'sm__inst_executed_pipe_', 'sm__pipe_fma', 'sm__pipe_alu', 'sm__pipe_xu',
        'smsp__issue_active.avg.pct', 'smsp__inst_executed.avg.per_cycle_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__sass_thread_inst_executed_op_f', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'smsp__warps_eligible.avg',
        'smsp__warp_issue_stalled', 'smsp__average_warp', 'local_load', 'local_store', 'l1tex__t_sector_hit_rate', 'lts__t_sector_hit_rate',
        'sm__sass_thread_inst_executed_op_', 'smsp__inst_issued.avg.per_cycle', 'smsp__pcsamp_warps_issue_stalled', 'derived__smsp__sass_thread_inst_executed_op']
for h, u, v in zip(hdr, units, vals):
    if any(w in h for w in want):
        print('%-100s %-14s %s' % (h, u, v))
