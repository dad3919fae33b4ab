"""Trace-driven replay of the render kernel's warp scheduling (render_v3.cuh) on the host.

Real paths of a scene are traced with the f64 oracle, every segment's op sequence is taken from the host-side walk of
the flattened stream (tests/opstream.py, `trace=`), and one warp's state machine - 32 lanes, class vote, slab / sphere
repetitions, shading quorum, lanes taking the next path of the warp's pool - is replayed over those sequences. Outputs
the quantities ncu measures for the real kernel (slab repetitions per path, lanes per slab repetition, lanes per
issued instruction, warp-instructions per path) so a scheduling idea can be judged before it costs GPU time; the
instruction costs per repetition come from the SASS view of the last capture (profiles/r1_render_v3_summary.md).

    python tools/warp_sim.py [--scene 8] [--tiles 12] [--samples 48] [--policy k=v,...]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rust_tracing_b200 as rt  # noqa: E402
from oracle import binding as ob  # noqa: E402
import opstream  # noqa: E402

SLAB, SPHERE, QUAD, MEDIUM, SHADE = 0, 1, 2, 3, 4
CLASS_OF_KIND = {0: SLAB, 1: SPHERE, 2: QUAD, 3: SLAB, 4: SLAB, 5: MEDIUM, 6: SLAB, 7: SLAB}
# warp-instructions per repetition / round (SASS view of capture r1_g; shade round fitted to the measured total)
COST = {"slab": 58.0, "slab_box_tail": 19.0, "slab_slow": 60.0, "sphere": 85.0, "quad": 75.0, "medium": 150.0,
        "vote_fast": 8.0, "vote_full": 40.0, "shade_round": 560.0}


def trace_paths(s, cam, tiles, samples, max_depth, seed=0):
    """Paths of `tiles` 8x4 tiles x `samples` samples in the order a warp's pool hands them out. Returns
    (segments: list over paths of lists of ray indices, rays)."""
    A = rt._abi
    h, w = cam.shape
    tiles_x = w // 8
    rng = np.random.default_rng(seed)
    tile_ids = rng.choice((w // 8) * (h // 4), size=tiles, replace=False)
    pix, smp = [], []
    for t in tile_ids:
        ty, tx = divmod(int(t), tiles_x)
        for sv in range(samples):
            for pv in range(32):
                pix.append((ty * 4 + (pv >> 3)) * w + tx * 8 + (pv & 7))
                smp.append(sv)
    pix, smp = np.array(pix), np.array(smp)
    n = len(pix)
    rays = ob.get_ray_batch(cam, pix, smp, seed=seed)
    lib = ob.lib()
    lib.oracle_scatter.restype = C.c_int
    path = np.arange(n)
    all_rays, seg_of_path = [], [[] for _ in range(n)]
    base = 0
    for seg in range(max_depth):
        if len(rays) == 0:
            break
        all_rays.append(rays.copy())
        for k, p in enumerate(path):
            seg_of_path[p].append(base + k)
        base += len(rays)
        hits = ob.hit_batch(s.desc, rays, seed=seed + 17 * seg)
        nxt = np.zeros(len(rays), dtype=A.ray_dtype())
        keep = np.zeros(len(rays), dtype=bool)
        sc = A.RayDesc()
        for k in np.flatnonzero(hits["hit"] == 1):
            r = A.RayDesc.from_buffer_copy(rays[k].tobytes())
            hd = A.HitDesc.from_buffer_copy(hits[k].tobytes())
            ok = lib.oracle_scatter(C.byref(s.desc), C.byref(r), C.byref(hd), C.c_uint64(seed), C.c_uint32(int(pix[path[k]])),
                                    C.c_uint32(int(smp[path[k]])), C.c_uint32(seg), 0, C.byref(sc), None, None)
            if ok == 1:
                nxt[k] = np.frombuffer(bytes(sc), dtype=A.ray_dtype())[0]
                keep[k] = True
        rays, path = nxt[keep], path[keep]
    return seg_of_path, np.concatenate(all_rays), tiles, samples


def op_sequences(S, rays):
    """Per ray: the op kinds it executes in the world program, in order."""
    trace = []
    opstream.hit_batch(S, rays, trace=trace)
    idx = np.concatenate([a for a, _ in trace])
    kind = np.concatenate([k for _, k in trace])
    it = np.concatenate([np.full(len(a), i) for i, (a, _) in enumerate(trace)])
    order = np.lexsort((it, idx))
    idx, kind = idx[order], kind[order]
    starts = np.searchsorted(idx, np.arange(len(rays)))
    ends = np.searchsorted(idx, np.arange(len(rays)), side="right")
    return [kind[a:b] for a, b in zip(starts, ends)]


def simulate(seg_of_path, seqs, tiles, samples, fold=True, shade_min=24, slab_fast=14, slab_reps=8, sphere_reps=2, slab_exit=1,
             postpone=0):
    """One warp per tile pool (pool = tile x all samples, as one launch chunk). Returns totals."""
    paths_per_pool = 32 * samples
    tot = {"instr": 0.0, "lane_instr": 0.0, "slab_reps": 0, "slab_lanes": 0, "sphere_reps": 0, "sphere_lanes": 0, "shade_rounds": 0,
           "shade_lanes": 0, "votes": 0, "paths": 0, "segments": 0}
    for pool in range(tiles):
        first = pool * paths_per_pool
        nxt_path = 0
        # lane state
        stash = [False] * 32      # postpone=1: the lane carries one sphere test it has not run yet
        ops = [None] * 32         # current segment's op kinds
        pos = [0] * 32
        segs = [None] * 32        # remaining segments (ray indices) of the lane's path
        cls = [SHADE] * 32        # SHADE = waiting (for a shade round / a new path); 5 = idle
        has = [False] * 32

        def lane_class(l):
            """Class the lane waits for. With postponed leaves a sphere op is stashed (the lane walks on through the
            stream) until a second sphere op or the end of the segment forces the stashed test to run."""
            while True:
                c = CLASS_OF_KIND[int(ops[l][pos[l]])] if pos[l] < len(ops[l]) else SHADE
                if not postpone:
                    return c
                if c == SPHERE:
                    if stash[l]:
                        return SPHERE            # flush first
                    stash[l] = True
                    pos[l] += 1
                    continue
                if c == SHADE and stash[l]:
                    return SPHERE                # the segment ends: run the stashed test
                return c

        while True:
            counts = [0] * 6
            for l in range(32):
                counts[cls[l]] += 1
            if counts[SLAB] >= slab_fast:
                pick = SLAB
                tot["instr"] += COST["vote_fast"]; tot["lane_instr"] += COST["vote_fast"] * 32
            else:
                if sum(counts[:5]) == 0:
                    break
                tot["instr"] += COST["vote_full"]; tot["lane_instr"] += COST["vote_full"] * 32
                if counts[SHADE] >= shade_min:
                    pick = SHADE
                else:
                    pick, best = SLAB, counts[SLAB]
                    for c in (SPHERE, QUAD, MEDIUM):
                        if counts[c] > best:
                            pick, best = c, counts[c]
                    if best == 0:
                        pick = SHADE
            tot["votes"] += 1
            if pick == SLAB:
                for rep in range(slab_reps):
                    lanes = [l for l in range(32) if cls[l] == SLAB]
                    if not lanes:
                        break
                    kinds = [int(ops[l][pos[l]]) for l in lanes]
                    n_box = sum(1 for k in kinds if k == 6)
                    n_slow = sum(1 for k in kinds if k in (3, 4, 7)) + (0 if fold else n_box)
                    cost = COST["slab"]
                    tot["lane_instr"] += COST["slab"] * len(lanes)
                    if fold and n_box:
                        cost += COST["slab_box_tail"]; tot["lane_instr"] += COST["slab_box_tail"] * n_box
                    if n_slow:
                        cost += COST["slab_slow"]; tot["lane_instr"] += COST["slab_slow"] * n_slow
                    tot["instr"] += cost
                    tot["slab_reps"] += 1; tot["slab_lanes"] += len(lanes)
                    for l in lanes:
                        pos[l] += 1
                        cls[l] = lane_class(l)
                    if sum(1 for l in range(32) if cls[l] == SLAB) < slab_exit:
                        break
            elif pick in (SPHERE, QUAD, MEDIUM):
                name = {SPHERE: "sphere", QUAD: "quad", MEDIUM: "medium"}[pick]
                for rep in range(sphere_reps if pick == SPHERE else 1):
                    lanes = [l for l in range(32) if cls[l] == pick]
                    if not lanes:
                        break
                    tot["instr"] += COST[name]; tot["lane_instr"] += COST[name] * len(lanes)
                    if pick == SPHERE:
                        tot["sphere_reps"] += 1; tot["sphere_lanes"] += len(lanes)
                    for l in lanes:
                        if postpone and pick == SPHERE:
                            if stash[l]:
                                stash[l] = False     # the stashed test ran; the lane's cursor already moved on
                            else:
                                pos[l] += 1
                        else:
                            pos[l] += 1
                        cls[l] = lane_class(l)
                    if postpone and pick == SPHERE:      # lanes of other classes flush their stash in the same repetition
                        for l in range(32):
                            if stash[l] and cls[l] != SPHERE and l not in lanes:
                                stash[l] = False
                                tot["lane_instr"] += COST[name]; tot["sphere_lanes"] += 1
            else:
                lanes = [l for l in range(32) if cls[l] == SHADE]
                tot["instr"] += COST["shade_round"]; tot["lane_instr"] += COST["shade_round"] * len(lanes)
                tot["shade_rounds"] += 1; tot["shade_lanes"] += len(lanes)
                for l in lanes:
                    if has[l] and segs[l]:
                        ray = segs[l].pop(0)               # the path goes on with its next segment
                    else:
                        if has[l]:
                            has[l] = False
                        if nxt_path < paths_per_pool:
                            p = first + nxt_path
                            nxt_path += 1
                            segs[l] = list(seg_of_path[p])
                            ray = segs[l].pop(0)
                            has[l] = True
                            tot["paths"] += 1
                        else:
                            cls[l] = 5                     # pool drained: idle
                            continue
                    ops[l], pos[l] = seqs[ray], 0
                    tot["segments"] += 1
                    cls[l] = lane_class(l)
    return tot


def report(tot, label):
    p = max(1, tot["paths"])
    print(f"{label}: warp-instr/path {tot['instr'] / p:7.1f}  lanes/instr {tot['lane_instr'] / max(tot['instr'], 1):5.2f}  "
          f"slab reps/path {tot['slab_reps'] / p:5.2f} at {tot['slab_lanes'] / max(1, tot['slab_reps']):5.2f} lanes  "
          f"sphere reps/path {tot['sphere_reps'] / p:5.2f} at {tot['sphere_lanes'] / max(1, tot['sphere_reps']):5.2f}  "
          f"shade rounds/path {tot['shade_rounds'] / p:5.3f} at {tot['shade_lanes'] / max(1, tot['shade_rounds']):5.2f}  "
          f"votes/path {tot['votes'] / p:5.2f}  segs/path {tot['segments'] / p:4.2f}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", type=int, default=8)
    ap.add_argument("--width", type=int, default=800)
    ap.add_argument("--tiles", type=int, default=12)
    ap.add_argument("--samples", type=int, default=48)
    ap.add_argument("--depth", type=int, default=40)
    a = ap.parse_args()
    earth, _ = rt.load_earth()
    s, cs = rt.builtin_scene(a.scene, image_width=a.width, max_depth=a.depth, earth=earth)
    cam = rt.Camera(cs)
    seg_of_path, rays, tiles, samples = trace_paths(s, cam, a.tiles, a.samples, a.depth)
    S = opstream.Stream(rt.scene_ops(s))
    seqs = op_sequences(S, rays)
    n_box = rt.scene_layout(s)["n_box"]
    print(f"scene {a.scene}: {len(seg_of_path)} paths, {len(rays)} segments, {sum(len(q) for q in seqs) / len(rays):.1f} stream ops per segment")
    fold = n_box >= 64
    report(simulate(seg_of_path, seqs, tiles, samples, fold=fold), "product policy            ")
    for kw in ({"shade_min": 16}, {"shade_min": 32}, {"slab_exit": 8}, {"slab_reps": 4}, {"slab_reps": 16}, {"slab_fast": 10}, {"slab_fast": 20},
               {"sphere_reps": 1}, {"sphere_reps": 4}, {"postpone": 1}, {"postpone": 1, "sphere_reps": 1}):
        report(simulate(seg_of_path, seqs, tiles, samples, fold=fold, **kw), f"{str(kw):26s}")


if __name__ == "__main__":
    main()
