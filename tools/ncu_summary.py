"""Print the handful of metrics we track from an .ncu-rep (run where ncu is installed; no GPU needed)."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__t_bytes.sum', 'lts__t_bytes.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed', 'sm__cycles_elapsed.avg.per_second',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'smsp__inst_executed_op_shared_ld.sum']


def summary(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    d = dict(zip(rows[0], rows[2]))
    u = dict(zip(rows[0], rows[1]))
    return {k: (d.get(k), u.get(k)) for k in KEYS}


if __name__ == '__main__':
    tabs = [summary(p) for p in sys.argv[1:]]
    for k in KEYS:
        print('%-90s' % k, '  '.join('%16s' % (t[k][0],) for t in tabs), tabs[0][k][1])
