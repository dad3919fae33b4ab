"""Host-side walk of the device's flattened traversal stream (rt_scene_ops_export), vectorised with numpy in f64.

Test infrastructure: it executes, op for op, what `traverse<>()` in csrc/device/rt_kernels.cuh executes on the
GPU - same skip links, same interval rules (sphere open, quad / box / medium closed), same instance folding, same
hoisted media, same keyed RNG - so the flattening done by scene_compile.cpp can be checked against the f64 oracle
without a GPU, and the number of ops a ray batch visits can be counted for stream-layout experiments
(tools/opstream_cost.py). It is not a fallback: nothing in the product imports it.
"""
import numpy as np

OP_INNER, OP_SPHERE, OP_QUAD, OP_XFORM_ENTER, OP_XFORM_EXIT, OP_MEDIUM, OP_BOX, OP_INNER_REF = range(8)
FLAG_MOVING, FLAG_PRECISE = 1, 2
MEDIUM_SPHERE, MEDIUM_PROGRAM, MEDIUM_XBOX = 0, 1, 2
P_MEDIUM = 16
KIND_NAMES = ["inner", "sphere", "quad", "xform_enter", "xform_exit", "medium", "box", "inner_ref"]


def pcg4d(v):
    """Jarzynski & Olano pcg4d on uint32 arrays of shape (n, 4) (rt_kernels.cuh:68-75)."""
    v = v.astype(np.uint64)
    M = np.uint64(0xFFFFFFFF)
    x, y, z, w = (v[:, k].copy() for k in range(4))
    x = (x * np.uint64(1664525) + np.uint64(1013904223)) & M
    y = (y * np.uint64(1664525) + np.uint64(1013904223)) & M
    z = (z * np.uint64(1664525) + np.uint64(1013904223)) & M
    w = (w * np.uint64(1664525) + np.uint64(1013904223)) & M

    def mix(x, y, z, w):
        x = (x + y * w) & M
        y = (y + z * x) & M
        z = (z + x * y) & M
        w = (w + y * z) & M
        return x, y, z, w
    x, y, z, w = mix(x, y, z, w)
    x ^= x >> np.uint64(16); y ^= y >> np.uint64(16); z ^= z >> np.uint64(16); w ^= w >> np.uint64(16)
    x, y, z, w = mix(x, y, z, w)
    return np.stack([x, y, z, w], axis=1).astype(np.uint32)


def path_key(seed, pixel, sample):
    n = len(pixel)
    v = np.zeros((n, 4), dtype=np.uint32)
    v[:, 0] = pixel
    v[:, 1] = sample
    v[:, 2] = np.uint32(seed & 0xFFFFFFFF)
    v[:, 3] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    return pcg4d(v)


def draw(key, seg, purpose):
    v = key.copy()
    v[:, 2] = (v[:, 2].astype(np.uint64) + np.uint64(seg)).astype(np.uint32)
    v[:, 3] = (v[:, 3].astype(np.uint64) + purpose.astype(np.uint64)).astype(np.uint32)
    return pcg4d(v)


def u01(x):
    return (x >> np.uint32(8)).astype(np.float64) * (1.0 / 16777216.0)


LINK_MASK = 0x0FFFFFFF
SLAB_SLACK = 1.0000012
OP_WORDS = {OP_INNER: 2, OP_SPHERE: 2, OP_QUAD: 4, OP_XFORM_ENTER: 4, OP_XFORM_EXIT: 2, OP_MEDIUM: 3, OP_BOX: 4, OP_INNER_REF: 2}


class Stream:
    """Header word: [0,8) size in bytes, [8,12) kind, [12,16) flags, [28,32) class of the fall-through successor.
    Skip links: byte offset | class << 28 (csrc/device/dev_scene.h)."""

    def __init__(self, ops):
        self.f = ops["words"].astype(np.float64)           # (N, 4) payloads
        self.i = ops["words"].view(np.int32).reshape(-1, 4)  # same words as integers (headers, links, ids)
        self.n_world = ops["n_world_words"]
        self.media = ops["media_ops"]
        self.first_link = ops.get("first_link", 0)

    def kind(self, idx):
        return (self.i[idx, 3] >> 8) & 15

    def flags(self, idx):
        return (self.i[idx, 3] >> 12) & 15

    def size_words(self, idx):
        return (self.i[idx, 3] & 0xFF) >> 4

    def skip(self, idx):
        """word index a skip link points to"""
        return ((self.i[idx + 1, 3].astype(np.int64) & LINK_MASK) >> 4)

    def walk(self):
        """(word index, kind, flags) of every op of the world program, in stream order."""
        out, i = [], 0
        while i < self.n_world:
            k, fl = int(self.kind(i)), int(self.flags(i))
            out.append((i, k, fl))
            i += int(self.size_words(i))
        return out


def _slab_ch(c, h, o, inv):
    """slab_ch() of rt_kernels.cuh: centre / half-extent form. Returns (te, tx, eps)."""
    with np.errstate(invalid="ignore", over="ignore"):
        oi = -o * inv
        tc = c * inv + oi
        th = np.abs(h * inv)
        near, far = tc - th, tc + th
    te = np.fmax(np.fmax(near[:, 0], near[:, 1]), near[:, 2])   # fmax / fmin drop a NaN operand like f64::max / min
    tx = np.fmin(np.fmin(far[:, 0], far[:, 1]), far[:, 2])
    e = np.abs(oi)
    e = np.where(np.isfinite(e), e, 0.0)
    return te, tx, 4.76837158e-7 * e.max(axis=1)


def _slab_interval(lo, hi, o, inv):
    with np.errstate(invalid="ignore", over="ignore"):
        a = (lo - o) * inv
        b = (hi - o) * inv
    neg = inv < 0.0
    near = np.where(neg, b, a)
    far = np.where(neg, a, b)
    te = np.fmax(np.fmax(near[:, 0], near[:, 1]), near[:, 2])   # fmax / fmin drop a NaN operand like f64::max / min
    tx = np.fmin(np.fmin(far[:, 0], far[:, 1]), far[:, 2])
    return te, tx, a, b


def _xform_point(x, w2, w3):
    q = x - w2[:, :3]
    s, c = w2[:, 3], w3[:, 3]
    return np.stack([c * q[:, 0] - s * q[:, 2] + w3[:, 0], q[:, 1] + w3[:, 1], s * q[:, 0] + c * q[:, 2] + w3[:, 2]], axis=1)


def _xform_dir(v, w2, w3):
    s, c = w2[:, 3], w3[:, 3]
    return np.stack([c * v[:, 0] - s * v[:, 2], v[:, 1], s * v[:, 0] + c * v[:, 2]], axis=1)


def _safe_inv(d):
    with np.errstate(divide="ignore"):
        return 1.0 / d


def _sphere_roots(oc, d, r):
    a = (d * d).sum(1)
    hb = (oc * d).sum(1)
    c = (oc * oc).sum(1) - r * r
    disc = hb * hb - a * c
    ok = disc >= 0.0
    sq = np.sqrt(np.where(ok, disc, 0.0))
    return ok, (-hb - sq) / a, (-hb + sq) / a


def traverse(S, o, d, time, tmin, tmax, begin=0, end=None, world=True, key=None, seg=0, counts=None, visits=None, trace=None):
    """Returns (t, op, xf): closest hit parameter, word index of the winning op (-1 = none) and of its enclosing
    OP_XFORM_ENTER (-1 = world space). `counts`: optional dict kind name -> visits, updated in place; `visits`:
    optional int64 array over stream words, incremented at the word index of every op a ray executes; `trace`: optional
    list that receives one (ray indices, op kinds) pair per lock-step iteration - per ray, in order, the ops it executes
    (tools/warp_sim.py replays them through the render kernel's warp scheduling)."""
    n = len(o)
    end = S.n_world if end is None else end
    o = o.astype(np.float64).copy(); d = d.astype(np.float64).copy()
    wo, wd = o.copy(), d.copy()                      # the world ray: instance ops hold composed world -> local transforms
    inv = _safe_inv(d)
    cur_xf = np.full(n, -1, dtype=np.int64)
    idx = np.full(n, begin, dtype=np.int64)
    best_t = np.full(n, tmax, dtype=np.float64) if np.isscalar(tmax) else tmax.astype(np.float64).copy()
    tmin = np.full(n, tmin, dtype=np.float64) if np.isscalar(tmin) else tmin
    best_op = np.full(n, -1, dtype=np.int64)
    best_xf = np.full(n, -1, dtype=np.int64)
    F, I = S.f, S.i

    def bump(name, m):
        if counts is not None:
            counts[name] = counts.get(name, 0) + int(m.sum())

    def medium(rows, at):
        """ConstantMedium::hit for rays `rows` (indices) whose medium op sits at word `at` (array). Returns next index."""
        w0f, w0i = F[at], I[at]
        bkind = (w0i[:, 3] >> 12) & 15
        nxt = np.zeros(len(rows), dtype=np.int64)
        t1 = np.full(len(rows), np.nan); t2 = np.full(len(rows), np.nan)
        ok = np.zeros(len(rows), dtype=bool)
        ro, rd = o[rows], d[rows]
        m = bkind == MEDIUM_SPHERE
        if m.any():
            a = at[m]
            w1, w2f, w2i = F[a + 1], F[a + 2], I[a + 2]
            moving = ((w2i[:, 3] >> 24) & FLAG_MOVING) != 0
            c = w1[:, :3] + np.where(moving[:, None], time[rows][m][:, None] * w2f[:, :3], 0.0)
            k, r1, r2 = _sphere_roots(ro[m] - c, rd[m], w1[:, 3])
            t1[m], t2[m] = r1, r2
            ok[m] = k & (r2 > r1 + 0.0001)
            nxt[m] = a + 3
        m = bkind == MEDIUM_XBOX
        if m.any():
            a = at[m]
            lo_ = _xform_point(ro[m], F[a + 1], F[a + 2]); ld_ = _xform_dir(rd[m], F[a + 1], F[a + 2])
            te, tx, _, _ = _slab_interval(F[a + 3][:, :3], F[a + 4][:, :3], lo_, _safe_inv(ld_))
            t1[m], t2[m] = te, tx
            ok[m] = (te <= tx) & (tx >= te + 0.0001) & np.isfinite(te) & np.isfinite(tx)
            nxt[m] = a + 5
        m = bkind == MEDIUM_PROGRAM
        if m.any():
            a = at[m]
            rr = rows[m]
            for b0, b1 in set(zip(I[a + 1][:, 0].tolist(), I[a + 1][:, 1].tolist())):
                sel = (I[a + 1][:, 0] == b0) & (I[a + 1][:, 1] == b1)
                r_ = rr[sel]
                ta, opa, _ = traverse(S, o[r_], d[r_], time[r_], -np.inf, np.inf, begin=b0, end=b1, world=False)
                hit1 = opa >= 0
                tb, opb, _ = traverse(S, o[r_], d[r_], time[r_], np.where(hit1, ta + 0.0001, np.inf), np.inf, begin=b0, end=b1, world=False)
                mm = np.flatnonzero(m)[sel]
                t1[mm], t2[mm] = ta, tb
                ok[mm] = hit1 & (opb >= 0)
                nxt[mm] = b1
        a1 = np.fmax(t1, tmin[rows]); a2 = np.fmin(t2, best_t[rows])
        ok &= a1 < a2
        a1 = np.fmax(a1, 0.0)
        length = np.sqrt((rd * rd).sum(1))
        with np.errstate(invalid="ignore"):
            inside = (a2 - a1) * length
        if ok.any():
            u = u01(draw(key[rows], seg, (P_MEDIUM + w0i[:, 2]).astype(np.uint32))[:, 0])
            with np.errstate(divide="ignore"):
                dist = w0f[:, 0] * np.log(u)
            win = ok & (dist <= inside)
            r_ = rows[win]
            best_t[r_] = (a1 + dist / length)[win]
            best_op[r_] = at[win]
            best_xf[r_] = cur_xf[r_]
        return nxt

    if world and S.media:
        rows = np.arange(n)
        for mop in S.media:
            bump("medium", np.ones(n, dtype=bool))
            medium(rows, np.full(n, mop, dtype=np.int64))

    for _ in range(10_000_000):
        act = np.flatnonzero(idx < end)
        if act.size == 0:
            break
        at = idx[act]
        kind = (I[at, 3] >> 8) & 15
        if visits is not None:
            np.add.at(visits, at, 1)
        if trace is not None:
            trace.append((act.copy(), kind.astype(np.int8)))
        # ---- box-headed ops: INNER, XFORM_ENTER, BOX (centre / half extent) and INNER_REF (the reference's corners)
        for k in (OP_INNER, OP_INNER_REF, OP_XFORM_ENTER, OP_BOX):
            m = kind == k
            if not m.any():
                continue
            bump(KIND_NAMES[k], m)
            r_, a = act[m], at[m]
            if k == OP_INNER_REF:                      # aabb.rs:64-84: per axis against the original interval
                _, _, pa, pb = _slab_interval(F[a][:, :3], F[a + 1][:, :3], o[r_], inv[r_])
                neg = inv[r_] < 0.0
                near = np.where(neg, pb, pa); far = np.where(neg, pa, pb)
                miss = (np.fmin(far, best_t[r_][:, None]) <= np.fmax(near, tmin[r_][:, None])).any(axis=1)
                idx[r_] = np.where(~miss, a + 2, S.skip(a))
                continue
            te, tx, eps = _slab_ch(F[a][:, :3], F[a + 1][:, :3], o[r_], inv[r_])
            if k == OP_BOX:                            # box_accept(): the primitive itself, no slack
                t = te.copy()
                bad = ~((tmin[r_] <= t) & (t <= best_t[r_]))
                t[bad] = tx[bad]
                win = (te <= tx) & (tmin[r_] <= t) & (t <= best_t[r_])
                w_ = r_[win]
                best_t[w_] = t[win]; best_op[w_] = a[win]; best_xf[w_] = cur_xf[w_]
                idx[r_] = a + 4
                continue
            passed = np.fmax(te, tmin[r_]) <= np.fmin(tx, best_t[r_]) * SLAB_SLACK + eps     # cull_pass()
            skip = S.skip(a)
            if k == OP_XFORM_ENTER:
                idx[r_[~passed]] = skip[~passed]
                p_ = r_[passed]; ap = a[passed]
                w2, w3 = F[ap + 2], F[ap + 3]
                o[p_] = _xform_point(wo[p_], w2, w3)
                d[p_] = _xform_dir(wd[p_], w2, w3)
                inv[p_] = _safe_inv(d[p_])
                cur_xf[p_] = ap
                idx[p_] = ap + 4
            else:
                idx[r_] = np.where(passed, a + 2, skip)
        m = kind == OP_XFORM_EXIT
        if m.any():
            r_ = act[m]
            parent = I[at[m], 0].astype(np.int64)          # the enclosing instance, -1 = world space
            o[r_] = wo[r_]; d[r_] = wd[r_]
            nested = parent >= 0
            if nested.any():
                q_, pp = r_[nested], parent[nested]
                o[q_] = _xform_point(wo[q_], F[pp + 2], F[pp + 3])
                d[q_] = _xform_dir(wd[q_], F[pp + 2], F[pp + 3])
            inv[r_] = _safe_inv(d[r_])
            cur_xf[r_] = parent
            idx[r_] = at[m] + 2
        m = kind == OP_SPHERE
        if m.any():
            bump("sphere", m)
            r_, a = act[m], at[m]
            flags = (I[a, 3] >> 12) & 15
            moving = (flags & FLAG_MOVING) != 0
            c = F[a][:, :3].copy()
            if moving.any():
                c[moving] += time[r_][moving][:, None] * F[a[moving] + 2][:, :3]
            ok, r1, r2 = _sphere_roots(o[r_] - c, d[r_], F[a + 1][:, 0])
            root = r1.copy()
            bad = ~((tmin[r_] < root) & (root < best_t[r_]))            # surrounds: open interval (sphere.rs:78-83)
            root[bad] = r2[bad]
            win = ok & (tmin[r_] < root) & (root < best_t[r_])
            w_ = r_[win]
            best_t[w_] = root[win]; best_op[w_] = a[win]; best_xf[w_] = cur_xf[w_]
            idx[r_] = a + np.where(moving, 3, 2)
        m = kind == OP_QUAD
        if m.any():
            bump("quad", m)
            r_, a = act[m], at[m]
            nrm = F[a][:, :3]
            denom = (nrm * d[r_]).sum(1)
            with np.errstate(divide="ignore", invalid="ignore"):
                t = (F[a + 3][:, 0] - (nrm * o[r_]).sum(1)) / denom
            p = o[r_] + t[:, None] * d[r_]
            alpha = (F[a + 1][:, :3] * p).sum(1) + F[a + 1][:, 3]
            beta = (F[a + 2][:, :3] * p).sum(1) + F[a + 2][:, 3]
            win = ~(np.abs(denom) < 1e-8) & (tmin[r_] <= t) & (t <= best_t[r_]) & ~((alpha < 0) | (alpha > 1) | (beta < 0) | (beta > 1))
            w_ = r_[win]
            best_t[w_] = t[win]; best_op[w_] = a[win]; best_xf[w_] = cur_xf[w_]
            idx[r_] = a + 4
        m = kind == OP_MEDIUM
        if m.any():
            r_, a = act[m], at[m]
            if world:
                bump("medium", m)
                idx[r_] = medium(r_, a)
            else:
                idx[r_] = end
    return best_t, best_op, best_xf


def hit_batch(S, rays, t_min=0.001, t_max=np.inf, seed=7, counts=None, visits=None, trace=None):
    """The emulated rt_hit_batch: structured array with hit / t / prim_id (the fields the flattening decides)."""
    o = np.ascontiguousarray(rays["origin"], dtype=np.float64)
    d = np.ascontiguousarray(rays["direction"], dtype=np.float64)
    time = np.ascontiguousarray(rays["time"], dtype=np.float64)
    n = len(o)
    key = path_key(seed, np.arange(n, dtype=np.uint32), np.zeros(n, dtype=np.uint32))
    t, op, xf = traverse(S, o, d, time, t_min, t_max, key=key, seg=0, counts=counts, visits=visits, trace=trace)
    out = np.zeros(n, dtype=[("hit", np.int32), ("t", np.float64), ("prim_id", np.int32)])
    hit = op >= 0
    out["hit"] = hit
    out["t"][hit] = t[hit]
    out["prim_id"] = -1
    F, I = S.f, S.i
    rows = np.flatnonzero(hit)
    a = op[rows]
    kind = (I[a, 3] >> 8) & 15
    pid = np.full(len(rows), -1, dtype=np.int64)
    m = kind == OP_SPHERE
    pid[m] = I[a[m] + 1, 2]
    m = kind == OP_QUAD
    pid[m] = I[a[m] + 3, 2]
    m = kind == OP_MEDIUM
    pid[m] = I[a[m], 2]
    m = kind == OP_BOX
    if m.any():
        r_, ab = rows[m], a[m]
        lo_o, lo_d = o[r_].copy(), d[r_].copy()
        inx = xf[r_] >= 0
        if inx.any():
            x = xf[r_][inx]
            lo_o[inx] = _xform_point(lo_o[inx], F[x + 2], F[x + 3])
            lo_d[inx] = _xform_dir(lo_d[inx], F[x + 2], F[x + 3])
        _, _, pa, pb = _slab_interval(F[ab + 2][:, :3], F[ab + 3][:, :3], lo_o, _safe_inv(lo_d))   # the exact corners
        tt = t[r_]
        # finalize_hit: the face whose plane parameter is nearest to the traversal's t; later faces of quad.rs:45-93 win ties
        face = np.zeros(len(r_), dtype=np.int64)
        miss = np.full(len(r_), np.inf)
        for f, plane in ((0, pb[:, 2]), (1, pb[:, 0]), (2, pa[:, 2]), (3, pa[:, 0]), (4, pb[:, 1]), (5, pa[:, 1])):
            with np.errstate(invalid="ignore"):
                mm = np.abs(plane - tt)
            upd = mm <= miss
            face[upd] = f
            miss[upd] = mm[upd]
            out["t"][r_[upd]] = plane[upd]
        pid[m] = I[ab + 2, 3] + face
    out["prim_id"][rows] = pid
    return out
