"""Turns the .ncu-rep captures in gpurun_out/ into the committed summaries under profiles/ (run where ncu exists)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPS = [('r1_a_v1_whole_segment_loop', 'prof_r1_a.ncu-rep', 'v1: every lane traces its whole segment; cube = 6 quads; media in the BVH'),
        ('r1_b_v2_class_voting', 'prof_r1_b.ncu-rep', 'v2: class-voting state machine + cube slab primitive'),
        ('r1_c4_v3_128regs', 'prof_r1_c4.ncu-rep', 'v3 @ 4 blocks/SM (cold state in shared memory), before the code-size work'),
        ('r1_c6_v3_80regs', 'prof_r1_c6.ncu-rep', 'v3 @ 6 blocks/SM, before the code-size work (I-cache thrash)'),
        ('r1_d_v3_small_code', 'prof_r1_d.ncu-rep', 'v3 @ 5 blocks/SM after shrinking the instruction footprint'),
        ('r1_e_v3_default', 'prof_r1_e.ncu-rep', 'v3 @ 6 blocks/SM, fast div/sqrt build'),
        ('r1_f_v3_pruned_stream', 'prof_r1_f.ncu-rep', 'v3 @ 6 blocks/SM on the pruned op stream (prune_stream)'),
        ('r1_g_v3_fold_box', 'prof_r1_g.ncu-rep', 'v3 FOLD form @ 6 blocks/SM on the pruned stream (current default for final_scene)')]
KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed', 'sm__cycles_elapsed.avg.per_second']
PATHS = 800 * 800 * 32


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))


def main():
    table, units = {}, {}
    reps = [r for r in REPS if os.path.exists(os.path.join(ROOT, 'gpurun_out', r[1]))]
    for name, fn, _ in reps:
        d, u = raw(os.path.join(ROOT, 'gpurun_out', fn))
        table[name] = {k: d.get(k) for k in KEYS}
        units.update({k: u.get(k) for k in KEYS})
    with open(os.path.join(ROOT, 'profiles', 'r1_render_kernel_ncu.csv'), 'w') as f:
        w = csv.writer(f)
        w.writerow(['metric', 'unit'] + [r[0] for r in reps])
        for k in KEYS:
            w.writerow([k, units[k]] + [table[r[0]][k] for r in reps])
    g = lambda t, k: float(t[k])
    L = []
    L.append('# Round 1 - ncu summaries of the render kernel (K1)\n')
    L.append('Command for every capture (one GPU, after the same command exited 0 without ncu):\n')
    L.append('    ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 \\')
    L.append('        python tools/ncu_target.py --spp 32 --reps 2      # final_scene 800x800, 32 spp, depth 40 = 20.48 M paths per launch\n')
    L.append('Raw metric table: `r1_render_kernel_ncu.csv`. Times under ncu are not bench values.\n')
    L.append('| capture | what | ms | regs | lanes/inst | warp-inst per path | issue active % | no_instruction | wait | L1 hit % | DRAM MB |')
    L.append('|---|---|---|---|---|---|---|---|---|---|---|')
    for name, _, desc in reps:
        t = table[name]
        L.append('| {} | {} | {:.1f} | {} | {} | {:.0f} | {:.1f} | {:.2f} | {:.2f} | {:.1f} | {:.0f} |'.format(
            name, desc, g(t, 'gpu__time_duration.sum'), t['launch__registers_per_thread'],
            t['smsp__thread_inst_executed_per_inst_executed.ratio'], g(t, 'smsp__inst_executed.sum') / PATHS,
            g(t, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
            g(t, 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio'),
            g(t, 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio'), g(t, 'l1tex__t_sector_hit_rate.pct'),
            g(t, 'dram__bytes_read.sum') + g(t, 'dram__bytes_write.sum')))
    last = reps[-1][0]
    d = table[last]
    hw = (g(d, 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed') + g(d, 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed')
          + 2 * g(d, 'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed'))
    L.append('\n## Reading of the current kernel ({})\n'.format(last))
    L.append('* Hardware view of FP32 work: (fadd + fmul + 2 ffma) = {:.0f} flop/cycle chip-wide = {:.1f}% of the 148 x 128 x 2 = 37 888 flop/cycle peak. '
             'The path is not FP32-throughput bound; it is bound by SIMT divergence and per-op overhead.'.format(hw, 100 * hw / 37888))
    L.append('* Lanes per executed instruction {} of 32: lanes of minority classes (sphere, shade) wait for the vote while the slab class runs.'.format(
        d['smsp__thread_inst_executed_per_inst_executed.ratio']))
    L.append('* {:.0f} warp-instructions per path; about half are slab-class repetitions (57 instructions per repetition including loop control and the next-op fetch, ~10-11 active lanes; SASS view of r1_e and of the wavefront extend kernel).'.format(
        g(d, 'smsp__inst_executed.sum') / PATHS))
    L.append('* Issue slots {:.1f}% busy; per issued instruction {:.2f} `wait` (fixed-latency dependency), {:.2f} `not_selected`, {:.2f} `no_instruction`, {:.2f} `branch_resolving`, {:.2f} `long_scoreboard`.'.format(
        g(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active'), g(d, 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio'),
        g(d, 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio'), g(d, 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio'),
        g(d, 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio'), g(d, 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio')))
    L.append('* I-cache: `no_instruction` was 3.26 (v3, 4 blocks) and 6.55 (6 blocks) before the instruction-footprint work and 0.90 right after it (r1_d); that change alone took the launch from 62.1 ms to 43.4 ms.')
    dram = (g(d, 'dram__bytes_read.sum') + g(d, 'dram__bytes_write.sum')) * 1e6
    L.append('* Memory: L1 hit {:.1f}%, L2 hit {:.1f}%, DRAM traffic {:.0f} MB per launch = {:.1f} B per path (earth texels and the framebuffer reductions); HBM is idle.'.format(
        g(d, 'l1tex__t_sector_hit_rate.pct'), g(d, 'lts__t_sector_hit_rate.pct'), dram / 1e6, dram / PATHS))
    L.append('\nPer-instruction (SASS) views used for these readings were exported with `ncu -i <rep> --page source --csv`; the `.ncu-rep` files stay in `gpurun_out/` (scratch, not committed).')
    with open(os.path.join(ROOT, 'profiles', 'r1_render_v3_summary.md'), 'w') as f:
        f.write('\n'.join(L) + '\n')
    json.dump({'kernel': 'render_kernel_v3', 'capture': last + ' (final_scene 800x800, 32 spp = 20.48 M paths per launch)',
               'dram_bytes_per_launch': dram, 'dram_bytes_per_path': dram / PATHS,
               'note': 'ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of ONE 20.48 M-path launch; the bench launch is 312x longer and its traffic scales per path'},
              open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json'), 'w'), indent=1)
    print('\n'.join(L))


if __name__ == '__main__':
    main()
