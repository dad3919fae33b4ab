"""Same-GPU A/B of whole library builds: every .so given on the command line renders the same scene, in alternating
order, timed with CUDA events on the launching stream (torch). Uses only the entry points every ABI version has
(rt_scene_builtin, rt_camera_new, rt_context_create, rt_scene_upload, rt_render_accumulate), so the round-1 library
(`git archive <round-1 commit>` built into csrc/librt_b200_r1.so) can stand next to the current one.

    python tools/ab_lib.py --scene 8 --spp 1000 --rounds 3 rust-tracing_b200/csrc/librt_b200_r1.so rust-tracing_b200/csrc/librt_b200.so
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--scene", type=int, default=8)
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--spp", type=int, default=1000)
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--env", action="append", default=[], help="NAME=VALUE set while the LAST library creates its context (RT_B200_DEV builds)")
    a = ap.parse_args()
    import torch
    import rust_tracing_b200 as rt      # structures only; each library is loaded separately below
    A = rt._abi
    earth, src = rt.load_earth()
    earth = np.ascontiguousarray(earth)
    runs = []
    for k, path in enumerate(a.libs):
        lib = C.CDLL(os.path.abspath(path))
        vp = C.c_void_p
        lib.rt_scene_builtin.argtypes = [C.POINTER(A.SceneRequest), C.POINTER(vp), C.POINTER(A.SceneDesc), C.POINTER(A.CameraSettingsC)]
        lib.rt_camera_new.argtypes = [C.POINTER(A.CameraSettingsC), C.POINTER(A.CameraDesc)]
        lib.rt_context_create.argtypes = [C.c_int, C.POINTER(vp)]
        lib.rt_scene_upload.argtypes = [vp, C.POINTER(A.SceneDesc), C.POINTER(vp)]
        lib.rt_render_accumulate.argtypes = [vp, vp, C.POINTER(A.CameraDesc), C.c_int64, C.c_int64, C.c_uint64, vp, vp]
        lib.rt_last_error.restype = C.c_char_p
        req = A.SceneRequest()
        req.scene, req.image_width, req.max_depth = a.scene, a.width, a.depth
        req.scene_seed, req.bvh_seed, req.perlin_seed = 1, 2, 3
        req.earth_height, req.earth_width = earth.shape[:2]
        req.earth_rgb8 = earth.ctypes.data_as(C.POINTER(C.c_uint8))
        raw, desc, cs, cam, ctx, ds = vp(), A.SceneDesc(), A.CameraSettingsC(), A.CameraDesc(), vp(), vp()

        def chk(rc, what):
            if rc < 0:
                raise RuntimeError(f"{path}: {what}: {lib.rt_last_error().decode()}")
        chk(lib.rt_scene_builtin(C.byref(req), C.byref(raw), C.byref(desc), C.byref(cs)), "scene")
        chk(lib.rt_camera_new(C.byref(cs), C.byref(cam)), "camera")
        if k == len(a.libs) - 1:
            for kv in a.env:
                n, v = kv.split("=", 1)
                os.environ[n] = v
        chk(lib.rt_context_create(0, C.byref(ctx)), "context")
        chk(lib.rt_scene_upload(ctx, C.byref(desc), C.byref(ds)), "upload")
        runs.append((path, lib, cam, ctx, ds, chk))
    h, w = int(runs[0][2].image_height), int(runs[0][2].image_width)
    fb = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda:0")
    stream = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    print(f"scene {a.scene} {w}x{h} spp {a.spp} depth {runs[0][2].max_depth} earth {src}", flush=True)
    res = {p: [] for p, *_ in runs}
    for rnd in range(-1, a.rounds):
        for path, lib, cam, ctx, ds, chk in (runs if rnd % 2 == 0 else runs[::-1]):
            fb.zero_()
            spp = a.spp if rnd >= 0 else max(1, a.spp // 8)      # round -1: warm-up
            e0.record(stream)
            chk(lib.rt_render_accumulate(ctx, ds, C.byref(cam), 0, spp, rnd + 7, fb.data_ptr(), stream.cuda_stream), "render")
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            assert float(fb[..., 3].min()) == spp
            if rnd >= 0:
                res[path].append(h * w * spp / ms / 1e3)
                print(f"  round {rnd} {os.path.basename(path):28s} {res[path][-1]:8.1f} Mpaths/s  mean {float(fb[..., :3].mean()) / spp:.5f}", flush=True)
    base = np.median(res[runs[0][0]])
    for p, *_ in runs:
        print(f"{os.path.basename(p):28s} median {np.median(res[p]):8.1f} Mpaths/s  x{np.median(res[p]) / base:.3f}")


if __name__ == "__main__":
    main()
