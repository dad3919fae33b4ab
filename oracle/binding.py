"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import
this module (see the header of oracle.cpp). The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")

COUNTER_NAMES = None
_lib = None


def build(force=False):
    src = os.path.join(HERE, "oracle.cpp")
    hdr = os.path.join(HERE, "..", "include", "rt_b200.h")
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return LIB_PATH
    subprocess.run(["make", "-C", HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return LIB_PATH


def lib():
    global _lib, COUNTER_NAMES
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        h = C.CDLL(LIB_PATH)
        vp, P = C.c_void_p, C.POINTER
        h.oracle_render.restype = C.c_int
        h.oracle_render.argtypes = [vp, vp, C.c_int64, C.c_int64, C.c_uint64, C.c_int, C.c_int, vp, vp, vp]
        h.oracle_hit_batch.restype = C.c_int
        h.oracle_hit_batch.argtypes = [vp, vp, C.c_int64, C.c_double, C.c_double, C.c_uint64, vp]
        h.oracle_texture_batch.restype = C.c_int
        h.oracle_texture_batch.argtypes = [vp, C.c_int, vp, C.c_int64, vp]
        h.oracle_get_ray_batch.restype = C.c_int
        h.oracle_get_ray_batch.argtypes = [vp, vp, vp, C.c_int64, C.c_uint64, vp]
        h.oracle_scatter.restype = C.c_int
        h.oracle_scatter.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, vp, vp, vp]
        h.oracle_camera_new.restype = C.c_int
        h.oracle_camera_new.argtypes = [vp, vp]
        h.oracle_bvh_build.restype = C.c_int
        h.oracle_bvh_build.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, P(C.c_int32), P(C.c_int32)]
        h.oracle_validate_scene.restype = C.c_int
        h.oracle_validate_scene.argtypes = [vp, P(C.c_double), P(C.c_double)]
        h.oracle_sphere_uv.argtypes = [vp, P(C.c_double), P(C.c_double)]
        h.oracle_reflectance.restype = C.c_double
        h.oracle_reflectance.argtypes = [C.c_double, C.c_double]
        h.oracle_refract.argtypes = [vp, vp, C.c_double, vp]
        h.oracle_reflect.argtypes = [vp, vp, vp]
        h.oracle_aabb_hit.restype = C.c_int
        h.oracle_aabb_hit.argtypes = [vp, vp, C.c_double, C.c_double]
        h.oracle_rgb_to_color.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8, vp]
        h.oracle_color_to_rgb.argtypes = [vp, vp]
        h.oracle_finalize_rgb8.argtypes = [vp, C.c_int64, C.c_double, vp]
        h.oracle_perlin_noise.restype = C.c_double
        h.oracle_perlin_noise.argtypes = [vp, vp]
        h.oracle_perlin_turbulence.restype = C.c_double
        h.oracle_perlin_turbulence.argtypes = [vp, vp, C.c_int]
        h.oracle_sample.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int64, vp]
        h.oracle_pcg4d.argtypes = [vp, vp]
        h.oracle_num_counters.restype = C.c_int
        h.oracle_counter_names.restype = C.c_char_p
        COUNTER_NAMES = h.oracle_counter_names().decode().split(",")
        assert len(COUNTER_NAMES) == h.oracle_num_counters()
        _lib = h
    return _lib


def _ptr(x):
    return C.addressof(x) if isinstance(x, C.Structure) else x.ctypes.data


def render(scene_desc, cam, sample_begin=0, sample_count=None, seed=0, mode=0, threads=0, want_sumsq=False):
    """renderer.rs:26-49 on the CPU in f64. Returns (sum_rgb (H,W,3) f64, counters dict[, sumsq_lum (H,W)])."""
    L = lib()
    if sample_count is None:
        sample_count = cam.samples_per_pixel
    h, w = int(cam.image_height), int(cam.image_width)
    out = np.zeros((h, w, 3), dtype=np.float64)
    sq = np.zeros((h, w), dtype=np.float64) if want_sumsq else None
    cnt = np.zeros(len(COUNTER_NAMES), dtype=np.uint64)
    rc = L.oracle_render(_ptr(scene_desc), _ptr(cam), sample_begin, sample_count, seed, mode, threads,
                         out.ctypes.data, sq.ctypes.data if want_sumsq else None, cnt.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle_render failed: {rc}")
    counters = dict(zip(COUNTER_NAMES, (int(x) for x in cnt)))
    return (out, counters, sq) if want_sumsq else (out, counters)


def hit_batch(scene_desc, rays, t_min=0.001, t_max=float("inf"), seed=7):
    from importlib import import_module
    A = import_module("rust_tracing_b200._abi")
    rays = np.ascontiguousarray(rays, dtype=A.ray_dtype())
    out = np.zeros(len(rays), dtype=A.hit_dtype())
    rc = lib().oracle_hit_batch(_ptr(scene_desc), rays.ctypes.data, len(rays), t_min, t_max, seed, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle_hit_batch failed: {rc}")
    return out


def texture_batch(scene_desc, tex, uvp):
    uvp = np.ascontiguousarray(uvp, dtype=np.float64).reshape(-1, 5)
    out = np.empty((len(uvp), 3), dtype=np.float64)
    rc = lib().oracle_texture_batch(_ptr(scene_desc), int(tex), uvp.ctypes.data, len(uvp), out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle_texture_batch failed: {rc}")
    return out


def get_ray_batch(cam, pixel_index, sample_index, seed=0):
    from importlib import import_module
    A = import_module("rust_tracing_b200._abi")
    pix = np.ascontiguousarray(pixel_index, dtype=np.int64)
    smp = np.ascontiguousarray(sample_index, dtype=np.int64)
    out = np.zeros(len(pix), dtype=A.ray_dtype())
    rc = lib().oracle_get_ray_batch(_ptr(cam), pix.ctypes.data, smp.ctypes.data, len(pix), seed, out.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle_get_ray_batch failed")
    return out


def camera_new(settings):
    from importlib import import_module
    A = import_module("rust_tracing_b200._abi")
    cam = A.CameraDesc()
    if lib().oracle_camera_new(C.addressof(settings), C.addressof(cam)) != 0:
        raise RuntimeError("oracle_camera_new failed")
    return cam


def bvh_build(boxes, axes):
    """bvh.rs:31-66 on bare boxes (n,6) with the given axis draws. Returns dict of arrays."""
    boxes = np.ascontiguousarray(boxes, dtype=np.float64).reshape(-1, 6)
    axes = np.ascontiguousarray(axes, dtype=np.int32)
    n = len(boxes)
    m = 2 * n - 1
    left = np.empty(m, np.int32); right = np.empty(m, np.int32); obj = np.empty(m, np.int32)
    bb = np.empty((m, 6), np.float64)
    nn, used = C.c_int32(), C.c_int32()
    rc = lib().oracle_bvh_build(boxes.ctypes.data, n, axes.ctypes.data, len(axes), left.ctypes.data, right.ctypes.data,
                                obj.ctypes.data, bb.ctypes.data, C.byref(nn), C.byref(used))
    if rc != 0:
        raise RuntimeError(f"oracle_bvh_build failed: {rc}")
    assert nn.value == m
    return {"left": left, "right": right, "object": obj, "bbox": bb, "axes_used": used.value}


def validate_scene(scene_desc):
    a, b = C.c_double(), C.c_double()
    if lib().oracle_validate_scene(_ptr(scene_desc), C.byref(a), C.byref(b)) != 0:
        raise RuntimeError("oracle_validate_scene failed")
    return a.value, b.value


def finalize_rgb8(sum_rgb, spp):
    s = np.ascontiguousarray(sum_rgb, dtype=np.float64).reshape(-1, 3)
    out = np.empty((len(s), 3), dtype=np.uint8)
    lib().oracle_finalize_rgb8(s.ctypes.data, len(s), float(spp), out.ctypes.data)
    return out


def sample(kind, mode, seed, n):
    out = np.empty((n, 3), dtype=np.float64)
    lib().oracle_sample(kind, mode, seed, n, out.ctypes.data)
    return out


def scatter_batch(scene_desc, rays, hits, pixel, sample, segment=0, seed=0, mode=0):
    """oracle_scatter over a batch: same structured layout as rt_scatter_batch (Material::emitted + scatter, material.rs:26-138)."""
    import rust_tracing_b200._abi as A
    rays = np.ascontiguousarray(rays, dtype=A.ray_dtype())
    hits = np.ascontiguousarray(hits, dtype=A.hit_dtype())
    out = np.zeros(len(rays), dtype=A.scatter_dtype())
    h = lib()
    att = (C.c_double * 3)()
    emi = (C.c_double * 3)()
    sc = A.RayDesc()
    for k in range(len(rays)):
        ok = h.oracle_scatter(C.byref(scene_desc), rays[k:k + 1].ctypes.data, hits[k:k + 1].ctypes.data, C.c_uint64(seed), int(pixel[k]),
                              int(sample[k]), int(segment), mode, C.byref(sc), att, emi)
        out["scattered"][k] = ok
        out["emitted"][k] = emi[:]
        if ok == 1:
            out["ray_out"][k] = np.frombuffer(bytes(sc), dtype=A.ray_dtype())[0]
            out["attenuation"][k] = att[:]
    return out
