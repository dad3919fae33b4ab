"""Region breakdown of a render-kernel ncu capture (ncu --set full --import-source on): issued warp-instructions per path
by source region, lanes per instruction, stall samples, plus the headline counters.

    python tools/ncu_regions.py gpurun_out/prof_r2_a.ncu-rep --paths 20480000 [--md profiles/r2_k1_region_breakdown.md]

Regions are (file, function) for the device functions of rt_kernels.cuh and labelled line ranges of the render loop
(render_mk.cuh / render_v3.cuh for round-1 captures, whose source is embedded in the report). Inlined code is attributed
to the function whose source line it came from (ncu's per-line view), so a helper shared by two callers is listed once.
"""
import argparse
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration ms",
    "launch__registers_per_thread": "registers / thread",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes per executed instruction (of 32)",
    "sm__inst_executed.sum": "warp instructions executed",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue slots busy %",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed": "ALU pipe busy %",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed": "FMA pipe busy %",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed": "FP64 pipe busy %",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed": "LSU instructions % of peak",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed": "XU instructions % of peak",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "L1/shared data pipe wavefronts % of peak",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared-memory wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared-memory bank conflicts",
    "smsp__inst_executed_op_shared_ld.sum": "shared loads (warp instructions)",
    "l1tex__t_sector_hit_rate.pct": "L1 hit %",
    "lts__t_sector_hit_rate.pct": "L2 hit %",
    "dram__bytes_read.sum": "DRAM bytes read",
    "dram__bytes_write.sum": "DRAM bytes written",
    "sm__warps_active.avg.per_cycle_active": "warps active per SM",
    "smsp__thread_inst_executed_pipe_fma.sum": None,
}
STALLS = ["wait", "long_scoreboard", "short_scoreboard", "not_selected", "branch_resolving", "no_instruction", "math_pipe_throttle",
          "mio_throttle", "lg_throttle", "barrier", "dispatch_stall", "selected"]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def raw_metrics(rep):
    rows = ncu_csv(rep, "raw")
    names, units, vals = rows[0], rows[1], rows[2]
    return {n: (v, u) for n, u, v in zip(names, units, vals)}


def function_table(src_text):
    """line -> name of the enclosing top-level device function (crude: a definition starts in column 0)."""
    lines = src_text.split("\n")
    out, cur = [None] * (len(lines) + 2), None
    for n, l in enumerate(lines, 1):
        m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__|inline|static).*?\b(\w+)\s*\(", l)
        if m and not l.startswith(" "):
            cur = m.group(1)
        out[n] = cur
    return out


# labelled line patterns of the render loop: first match wins, searched on the source TEXT of the line's region start
LOOP_MARKS = [
    ("vote", r"// ---- the vote|const unsigned n_slab = __popc|unsigned n_slab = __popc"),
    ("slab loop: rare kinds (cube accept, instance enter / exit)", r"^\s*if \(rare\) \{"),
    ("slab loop", r"if \(pick == CLS_SLAB\)|// ---- cull boxes|link = in_class \? nl : link"),
    ("sphere class", r"if \(pick == CLS_SPHERE\)"),
    ("quad class", r"else if \(pick == CLS_QUAD\)"),
    ("medium class", r"else if \(pick == CLS_MEDIUM\)"),
    ("box class", r"else if \(pick == CLS_BOX\)"),
    ("shade call / hand-over", r"// ---- shade"),
    ("entry into the box-test loop after another class", r"if \(n_slab < \(unsigned\)slab_fast\) continue"),
    ("epilogue", r"#undef CNT"),
]


def loop_regions(src_text):
    lines = src_text.split("\n")
    region, cur = [None] * (len(lines) + 2), "prologue"
    in_kernel = False
    for n, l in enumerate(lines, 1):
        if "__global__" in l:
            in_kernel, cur = True, "prologue"
        if in_kernel:
            for name, pat in LOOP_MARKS:
                if re.search(pat, l):
                    cur = name
                    break
        region[n] = cur if in_kernel else None
    return region


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--paths", type=float, required=True, help="paths traced by the captured launch")
    ap.add_argument("--md", default=None)
    ap.add_argument("--title", default=None)
    a = ap.parse_args()
    raw = raw_metrics(a.rep)
    rows = ncu_csv(a.rep, "source", ["--print-source", "sass,cuda"])
    # embedded sources: ncu prints them line by line only where instructions map, so fetch the full text through the cuda view
    hdr = None
    per = collections.OrderedDict()
    lanes = collections.Counter()
    samples = collections.Counter()
    cur_file = None
    files = {}
    for r in rows:
        if r and r[0] == "File Path":
            cur_file = r[1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            ci, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
            continue
        if not r or not r[0].isdigit() or hdr is None:
            continue
        try:
            n, c, t, s = int(r[0]), int(r[ci]), int(r[ti]), int(r[si])
        except (ValueError, IndexError):
            continue
        files.setdefault(cur_file, {})[n] = (r[1], c, t, s)
    total = sum(c for f in files.values() for (_, c, _, _) in f.values())
    total_s = sum(s for f in files.values() for (_, _, _, s) in f.values())
    agg = collections.Counter()
    for path, lines in files.items():
        base = path.split("/")[-1]
        try:
            text = open(path).read()
        except OSError:
            text = None
        ftab = function_table(text) if text and base.endswith(".cuh") and "render_" not in base else None
        ltab = loop_regions(text) if text and base.startswith("render_") else None
        for n, (src, c, t, s) in lines.items():
            if ltab and n < len(ltab) and ltab[n]:
                key = f"{base}: {ltab[n]}"
                if ltab[n] in ("prologue",) and text:
                    ft = function_table(text)
                    if ft[n] and ft[n] != "render_kernel_mk" and not ft[n].startswith("render_kernel"):
                        key = f"{base}: {ft[n]}()"
            elif ftab and n < len(ftab) and ftab[n]:
                key = f"{base}: {ftab[n]}()"
            else:
                key = base
            agg[key] += c
            lanes[key] += t
            samples[key] += s
    out = []
    w = out.append
    w(f"# {a.title or a.rep}")
    w("")
    w(f"Capture: `{a.rep}`, {a.paths / 1e6:.2f} M paths in the launch.")
    w("")
    w("| counter | value |")
    w("|---|---|")
    for k, label in KEYS.items():
        if label and k in raw:
            v, u = raw[k]
            try:
                fv = float(v)
                v = f"{fv:,.2f}" if abs(fv) < 1e6 else f"{fv:,.0f}"
            except ValueError:
                pass
            w(f"| {label} | {v} {u} |")
    if "sm__inst_executed.sum" in raw:
        w(f"| warp instructions per path | {float(raw['sm__inst_executed.sum'][0]) / a.paths:.1f} |")
    if "dram__bytes_read.sum" in raw:
        tot = float(raw["dram__bytes_read.sum"][0]) + float(raw["dram__bytes_write.sum"][0])
        unit = raw["dram__bytes_read.sum"][1]
        w(f"| DRAM traffic per launch | {tot:,.1f} {unit} |")
    for s_ in STALLS:
        k = f"smsp__average_warps_issue_stalled_{s_}_per_issue_active.ratio"
        if k in raw:
            w(f"| stall {s_} per issue | {float(raw[k][0]):.2f} |")
    w("")
    w("| region | warp-inst / path | % of issued | lanes / inst | % of stall samples |")
    w("|---|---|---|---|---|")
    for k, c in sorted(agg.items(), key=lambda kv: -kv[1]):
        if c < 0.002 * total:
            continue
        w(f"| {k} | {c / a.paths:.1f} | {100.0 * c / total:.1f} | {lanes[k] / max(c, 1):.1f} | {100.0 * samples[k] / max(total_s, 1):.1f} |")
    w(f"| (all) | {total / a.paths:.1f} | 100 | {sum(lanes.values()) / max(total, 1):.1f} | 100 |")
    text = "\n".join(out) + "\n"
    if a.md:
        open(a.md, "w").write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
