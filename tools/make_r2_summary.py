"""profiles/r2_render_kernel_ncu.csv: the raw ncu metrics of every round-2 capture of the render kernel, side by side
(same command for all: ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1
python tools/ncu_target.py --spp 32 --reps 2 = final_scene 800x800, 20.48 M paths per launch)."""
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_profile_summary import KEYS, raw, ROOT  # noqa: E402

REPS = [("r2_a_mk_first", "prof_r2_a.ncu-rep", "render_kernel_mk as first committed (one CTA per SM, op stream in shared memory)"),
        ("r2_b_mk_tuned", "prof_r2_b.ncu-rep", "after the vote / self-origin work (bench 1010)"),
        ("r2_c_mk_before_icache", "prof_r2_c.ncu-rep", "same kernel, final verification box (bench 946)"),
        ("r2_q1_two_queues", "prof_r2_q1.ncu-rep", "render_q.cuh, traverse / shade queues (not adopted)"),
        ("r2_q2_class_queues", "prof_r2_q2.ncu-rep", "render_q.cuh, one queue per op class, sticky roles (not adopted)"),
        ("r2_q3_class_queues", "prof_r2_q3.ncu-rep", "render_q.cuh, roles follow the fullest queue (not adopted)"),
        ("r2_d_mk_approx_div", "prof_r2_d.ncu-rep", "render_kernel_mk with .approx.ftz division / reciprocal / sqrt"),
        ("r2_e_mk_final", "prof_r2_e.ncu-rep", "final build of round 2 (bench 1191)")]


def main():
    reps = [r for r in REPS if os.path.exists(os.path.join(ROOT, "gpurun_out", r[1]))]
    table, units = {}, {}
    for name, fn, _ in reps:
        d, u = raw(os.path.join(ROOT, "gpurun_out", fn))
        table[name] = {k: d.get(k) for k in KEYS}
        units.update({k: u.get(k) for k in KEYS})
    with open(os.path.join(ROOT, "profiles", "r2_render_kernel_ncu.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [r[0] for r in reps])
        w.writerow(["what", ""] + [r[2] for r in reps])
        for k in KEYS:
            w.writerow([k, units[k]] + [table[r[0]][k] for r in reps])
    print("wrote", len(reps), "captures")


if __name__ == "__main__":
    main()
