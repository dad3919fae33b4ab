"""Offline study (no GPU): how many box tests would a different cull hierarchy over the SAME primitives cost?

The device stream visits primitives in the reference's depth-first order and uses the reference's BVH nodes (minus the
ones prune_stream drops) as cull boxes. Boxes only cull, so any hierarchy of conservative boxes gives the same hits:
  order   re-tree the leaf SEQUENCE top-down with a surface-area sweep (every node covers a contiguous run of the
          reference's leaf order: visiting order, hence every tie rule, untouched)
  free    a classic SAH tree over the leaves (changes the visiting order: ties would need the rank rule)
Counts ops per segment on the segments of real paths (tools/opstream_cost.py) with the host walk (tests/opstream.py).

    python tools/retree_study.py --scene 8 --paths 8000
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import rust_tracing_b200 as rt  # noqa: E402
import opstream  # noqa: E402
from opstream import OP_INNER, OP_SPHERE, OP_QUAD, OP_XFORM_ENTER, OP_XFORM_EXIT, OP_MEDIUM, OP_BOX, OP_INNER_REF  # noqa: E402
from opstream_cost import path_segments  # noqa: E402

LEAF_COST = {OP_SPHERE: 2.5, OP_QUAD: 1.5, OP_BOX: 1.5, OP_MEDIUM: 6.0, OP_XFORM_ENTER: 3.0}


def parse(S, desc, begin, end):
    """-> list of nodes: dict(kind, at, n_words, lo, hi, ch)"""
    out, i = [], begin
    F, I = S.f, S.i
    while i < end:
        k, fl = int(S.kind(i)), int(S.flags(i))
        n = {"kind": k, "at": i, "ch": []}
        if k in (OP_INNER, OP_INNER_REF):
            nxt = int(S.skip(i))
            n["n_words"] = 2
            n["ch"] = parse(S, desc, i + 2, nxt)
            if k == OP_INNER:
                n["lo"], n["hi"] = F[i, :3] - F[i + 1, :3], F[i, :3] + F[i + 1, :3]
            else:
                n["lo"], n["hi"] = F[i, :3].copy(), F[i + 1, :3].copy()
        elif k == OP_XFORM_ENTER:
            nxt = int(S.skip(i))
            n["n_words"] = 4
            n["ch"] = parse(S, desc, i + 4, nxt - 2)
            n["exit_at"] = nxt - 2
            n["lo"], n["hi"] = F[i, :3] - F[i + 1, :3], F[i, :3] + F[i + 1, :3]
        else:
            n["n_words"] = int(S.size_words(i))
            nxt = i + n["n_words"]
            if k == OP_SPHERE:
                c, r = F[i, :3], abs(F[i + 1, 0])
                lo, hi = c - r, c + r
                if fl & 1:
                    c2 = c + F[i + 2, :3]
                    lo, hi = np.minimum(lo, c2 - r), np.maximum(hi, c2 + r)
                n["lo"], n["hi"] = lo, hi
            elif k == OP_BOX:
                n["lo"], n["hi"] = F[i + 2, :3].copy(), F[i + 3, :3].copy()
            elif k == OP_QUAD:
                b = np.array(desc.hittables[int(I[i + 3, 2])].bbox[:])
                n["lo"], n["hi"] = b[0::2].copy(), b[1::2].copy()
            elif k == OP_MEDIUM:
                n["lo"], n["hi"] = None, None          # takes the box of the OP_INNER in front of it
                if fl == 1:
                    nxt = int(I[i + 1, 1])
                    n["n_words"] = nxt - i
        out.append(n)
        i = nxt
    return out


def leaves_of(nodes):
    """flatten cull boxes away; a medium keeps its own box node"""
    out = []
    for n in nodes:
        if n["kind"] == OP_INNER and not (len(n["ch"]) == 1 and n["ch"][0]["kind"] == OP_MEDIUM):
            out += leaves_of(n["ch"])
        else:
            out.append(n)
    return out


def area(lo, hi):
    e = np.maximum(hi - lo, 0.0)
    return 2.0 * (e[0] * e[1] + e[1] * e[2] + e[0] * e[2])


def cost_of(n):
    return 7.0 if n["kind"] == OP_INNER else LEAF_COST.get(n["kind"], 2.0)


def build_order(L):
    """top-down SAH sweep over the sequence L (contiguous runs only)."""
    if len(L) <= 1:
        return L
    n = len(L)
    lo = np.array([x["lo"] for x in L]); hi = np.array([x["hi"] for x in L])
    c = np.array([cost_of(x) for x in L])
    plo = np.minimum.accumulate(lo, axis=0); phi = np.maximum.accumulate(hi, axis=0)
    slo = np.minimum.accumulate(lo[::-1], axis=0)[::-1]; shi = np.maximum.accumulate(hi[::-1], axis=0)[::-1]
    cc = np.cumsum(c)
    best, bk = np.inf, -1
    for k in range(1, n):
        v = area(plo[k - 1], phi[k - 1]) * cc[k - 1] + area(slo[k], shi[k]) * (cc[-1] - cc[k - 1])
        if v < best:
            best, bk = v, k
    left, right = build_order(L[:bk]), build_order(L[bk:])
    return [wrap(left), wrap(right)]


def wrap(nodes):
    if len(nodes) == 1:
        return nodes[0]
    lo = np.min([x["lo"] for x in nodes], axis=0); hi = np.max([x["hi"] for x in nodes], axis=0)
    return {"kind": OP_INNER, "at": -1, "n_words": 2, "lo": lo, "hi": hi, "ch": nodes}


def build_free(L):
    if len(L) <= 1:
        return L
    n = len(L)
    lo = np.array([x["lo"] for x in L]); hi = np.array([x["hi"] for x in L])
    c = np.array([cost_of(x) for x in L])
    cen = 0.5 * (lo + hi)
    best, bsplit = np.inf, None
    for ax in range(3):
        o = np.argsort(cen[:, ax], kind="stable")
        l2, h2, c2 = lo[o], hi[o], c[o]
        plo = np.minimum.accumulate(l2, axis=0); phi = np.maximum.accumulate(h2, axis=0)
        slo = np.minimum.accumulate(l2[::-1], axis=0)[::-1]; shi = np.maximum.accumulate(h2[::-1], axis=0)[::-1]
        cc = np.cumsum(c2)
        e1 = np.maximum(phi - plo, 0); a1 = 2 * (e1[:, 0] * e1[:, 1] + e1[:, 1] * e1[:, 2] + e1[:, 0] * e1[:, 2])
        e2 = np.maximum(shi - slo, 0); a2 = 2 * (e2[:, 0] * e2[:, 1] + e2[:, 1] * e2[:, 2] + e2[:, 0] * e2[:, 2])
        v = a1[:-1] * cc[:-1] + a2[1:] * (cc[-1] - cc[:-1])
        k = int(np.argmin(v))
        if v[k] < best:
            best, bsplit = v[k], (o, k + 1)
    o, k = bsplit
    Ls = [L[j] for j in o]
    return [wrap(build_free(Ls[:k])), wrap(build_free(Ls[k:]))]


def prune(node, anc_area):
    """prune_stream's rule (scene_compile.cpp): returns (list of nodes to emit under an ancestor of area anc_area, cost)."""
    if node["kind"] != OP_INNER or node.get("keep"):
        if node["kind"] == OP_INNER:       # a medium's own box
            return [node], 1.0 + min(1.0, area(node["lo"], node["hi"]) / anc_area if np.isfinite(anc_area) and anc_area > 0 else 1.0) * 6.0
        return [node], cost_of(node) if node["kind"] != OP_INNER else 1.0
    a = area(node["lo"], node["hi"])
    p = min(1.0, a / anc_area) if np.isfinite(anc_area) and anc_area > 0 else 1.0
    kept, ck = [], 0.0
    for ch in node["ch"]:
        k, c = prune(ch, a)
        kept += k; ck += c
    keep_cost = 1.0 + p * ck
    dis, cd = [], 0.0
    for ch in node["ch"]:
        k, c = prune(ch, anc_area)
        dis += k; cd += c
    if keep_cost <= cd:
        m = dict(node); m["ch"] = kept
        return [m], keep_cost
    return dis, cd


def emit(S, nodes, words, xf_parent=-1):
    F = S.f
    raw = S.raw
    for n in nodes:
        k = n["kind"]
        if k == OP_INNER and n["at"] < 0 or (k == OP_INNER and "synth" in n):
            pass
        if k == OP_INNER:
            pos = len(words)
            c = 0.5 * (n["lo"] + n["hi"]); h = 0.5 * (n["hi"] - n["lo"])
            h = h + (np.abs(c) + h) * 2.0 ** -21 + 1e-30
            w0 = np.zeros(4, dtype=np.float32); w1 = np.zeros(4, dtype=np.float32)
            w0[:3] = c; w1[:3] = h
            w0.view(np.uint32)[3] = 32          # size 2 words, kind 0
            words.append(w0); words.append(w1)
            emit(S, n["ch"], words, xf_parent)
            w1.view(np.uint32)[3] = len(words) << 4
        elif k == OP_INNER_REF:
            pos = len(words)
            words.append(raw[n["at"]].copy()); words.append(raw[n["at"] + 1].copy())
            emit(S, n["ch"], words, xf_parent)
            words[pos + 1].view(np.uint32)[3] = len(words) << 4
        elif k == OP_XFORM_ENTER:
            pos = len(words)
            for j in range(4):
                words.append(raw[n["at"] + j].copy())
            emit(S, n["ch"], words, pos)
            e0, e1 = raw[n["exit_at"]].copy(), raw[n["exit_at"] + 1].copy()
            e0.view(np.int32)[0] = xf_parent
            words.append(e0); words.append(e1)
            words[pos + 1].view(np.uint32)[3] = len(words) << 4
        else:
            for j in range(n["n_words"]):
                words.append(raw[n["at"] + j].copy())


def restream(S, tree):
    words = []
    emit(S, tree, words)
    n = len(words)
    tail = [S.raw[j].copy() for j in range(S.n_world, len(S.raw))]
    media = [n + (m - S.n_world) for m in S.media]
    W = np.array(words + tail, dtype=np.float32)
    return opstream.Stream({"words": W, "n_world_words": n, "media_ops": media, "first_link": 0})


def retree(nodes, how):
    """rebuild the hierarchy of one space; instances keep their place as leaves and are re-treed inside."""
    L = leaves_of(nodes)
    for x in L:
        if x["kind"] == OP_XFORM_ENTER:
            x["ch"] = retree(x["ch"], how)
        if x["kind"] == OP_INNER:               # medium box: keep as is
            x["keep"] = True
    if any(x["kind"] == OP_INNER_REF for x in L) or len(L) <= 2:
        return nodes
    t = build_order(L) if how == "order" else build_free(L)
    root = wrap(t)
    out, _ = prune(root, np.inf)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", type=int, default=8)
    ap.add_argument("--width", type=int, default=800)
    ap.add_argument("--paths", type=int, default=8000)
    ap.add_argument("--depth", type=int, default=40)
    a = ap.parse_args()
    sys.setrecursionlimit(10000)
    earth, _ = rt.load_earth()
    s, cs = rt.builtin_scene(a.scene, image_width=a.width, max_depth=a.depth, earth=earth)
    cam = rt.Camera(cs)
    rays, n_paths = path_segments(s, cam, a.paths, a.depth)
    print(f"scene {a.scene}: {n_paths} paths, {len(rays)} segments ({len(rays) / n_paths:.2f} per path)")

    def report(name, S):
        counts = {}
        h = opstream.hit_batch(S, rays, counts=counts)
        box_class = counts.get("inner", 0) + counts.get("box", 0) + counts.get("xform_enter", 0) + counts.get("inner_ref", 0)
        print(f"{name:28s} words {S.n_world:6d}  per segment: inner {counts.get('inner', 0) / len(rays):6.2f}  box {counts.get('box', 0) / len(rays):5.2f}"
              f"  sphere {counts.get('sphere', 0) / len(rays):5.2f}  quad {counts.get('quad', 0) / len(rays):5.2f}  slab-class {box_class / len(rays):6.2f}", flush=True)
        return h

    Sp = opstream.Stream(rt.scene_ops(s)); Sp.raw = rt.scene_ops(s)["words"]
    h0 = report("product (pruned reference)", Sp)
    ops_full = rt.scene_ops(s, rt.layout_flags(prune=False))
    Sf = opstream.Stream(ops_full); Sf.raw = ops_full["words"]
    report("reference tree, unpruned", Sf)
    for how in ("order", "free"):
        tree = retree(parse(Sf, s.desc, 0, Sf.n_world), how)
        S2 = restream(Sf, tree)
        h = report(f"re-tree '{how}' + prune", S2)
        same = (h["hit"] == h0["hit"]).mean(), (np.abs(h["t"] - h0["t"]) <= 1e-9 * np.maximum(1, np.abs(h0["t"]))).mean()
        print(f"    same hit/miss {same[0]:.5f}  same t {same[1]:.5f}")


if __name__ == "__main__":
    main()
