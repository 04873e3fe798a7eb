"""Print the handful of ncu raw-page metrics this project cares about.
usage: python tools/ncu_summary.py X.ncu-rep"""
import csv
import subprocess
import sys

KEYS = [
    'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__waves_per_multiprocessor',
    'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'smsp__thread_inst_executed_per_inst_executed.ratio',
]


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    names = [r[hdr.index('Kernel Name')][:50] for r in rows[2:]]
    print('kernels:', names)
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print('%-90s %-8s %s' % (k, units[i], [r[i] for r in rows[2:]]))


if __name__ == '__main__':
    main(sys.argv[1])
