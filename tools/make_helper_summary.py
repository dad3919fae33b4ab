"""profiles/r2_helper_kernels_ncu.md from an `ncu --set full` capture of tools/ncu_small_kernels.py:

    python tools/make_helper_summary.py gpurun_out/prof_r2_small.ncu-rep
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# algorithmic bytes per launch (what the kernel must read + write), by kernel name and grid size
EARTH = 6400 * 3200


def algo_bytes(name, grid):
    if "expand_image" in name:
        return EARTH * 19, "3 B RGB8 in + 16 B float4 out per texel"
    if "finalize" in name:
        return 800 * 800 * 19, "16 B sum in + 3 B RGB8 out per pixel"
    if "jpeg_idct" in name:
        blocks = grid * 128
        return blocks * (128 + 64), "per 8x8 block: 128 B coefficients in + 64 B samples out (grid x 128 blocks, incl. the ragged last CTA)"
    if "jpeg_colour" in name:
        return EARTH * (1 + 0.5 + 3), "per pixel: 1 B luma + 2 x 1/4 B chroma in, 3 B RGB out (4:2:0)"
    return None, "latency-sized (1000 objects): no bandwidth figure"


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    col = {n: i for i, n in enumerate(hdr)}
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    peak = float(peaks.get("hbm_gbs", 6538.3))
    L = ["# Round 2 - helper kernels (ncu --set full --clock-control none, `python tools/ncu_small_kernels.py`)", "",
         f"Peak = {peak:.1f} GB/s (MEASURED_PEAKS.json `hbm_gbs`, burst figure: these kernels are timed alone). Times under ncu are cold-cache.",
         "", "| kernel | grid x block | duration | algorithmic bytes | achieved GB/s (algorithmic) | frac of measured HBM peak | DRAM traffic (ncu read+write) | regs |",
         "|---|---|---|---|---|---|---|---|"]
    notes = {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        short = name.split("(")[0].split("::")[-1]
        grid = int(r[col["launch__grid_size"]])
        block = int(r[col["launch__block_size"]])
        ns = float(r[col["gpu__time_duration.sum"]])
        unit = rows[1][col["gpu__time_duration.sum"]]
        us = ns / 1e3 if unit in ("ns", "nsecond") else ns if unit in ("us", "usecond") else ns * 1e3
        rd, wr = float(r[col["dram__bytes_read.sum"]]), float(r[col["dram__bytes_write.sum"]])
        ru = rows[1][col["dram__bytes_read.sum"]]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(ru, 1)
        dram = (rd + wr) * scale
        ab, why = algo_bytes(short, grid)
        notes[short] = why
        regs = r[col["launch__registers_per_thread"]]
        if ab:
            gbs = ab / (us * 1e-6) / 1e9
            L.append(f"| {short} | {grid} x {block} | {us:.1f} us | {ab / 1e6:.1f} MB | {gbs:.0f} | {gbs / peak:.2f} | {dram / 1e6:.1f} MB | {regs} |")
        else:
            L.append(f"| {short} | {grid} x {block} | {us:.1f} us | - | - | - | {dram / 1e6:.2f} MB | {regs} |")
    L.append("")
    for k, v in notes.items():
        L.append(f"* `{k}`: {v}.")
    text = "\n".join(L) + "\n"
    open(os.path.join(ROOT, "profiles", "r2_helper_kernels_ncu.md"), "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
