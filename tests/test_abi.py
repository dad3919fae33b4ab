"""The C-ABI library loads on a machine without a GPU and exports every symbol include/rt_b200.h declares."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_declares_the_boundary():
    names = declared_symbols()
    for must in ("rt_render", "rt_render_accumulate", "rt_scene_upload", "rt_hit_batch", "rt_camera_new",
                 "rt_hit_sphere", "rt_hit_quad", "rt_hit_cube", "rt_hit_bvh", "rt_hit_constant_medium",
                 "rt_scene_builtin", "rt_finalize_rgb8", "rt_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(rt):
    lib = C.CDLL(rt._abi.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"declared in rt_b200.h but not exported: {missing}"
    # and the Python prototypes cover the same set
    assert sorted(rt._abi.PROTOTYPES) == declared_symbols()


def test_abi_version(rt):
    assert rt._abi.lib().rt_abi_version() == 2


def test_struct_sizes_match_header(rt):
    A = rt._abi
    assert C.sizeof(A.TextureDesc) == 48
    assert C.sizeof(A.MaterialDesc) == 40
    assert C.sizeof(A.HittableDesc) == 24 + 15 * 8 + 2 * 8 + 6 * 8
    assert C.sizeof(A.BvhNodeDesc) == 64
    assert C.sizeof(A.PerlinDesc) == 256 * 24 + 3 * 1024
    assert C.sizeof(A.RayDesc) == 56 and A.ray_dtype().itemsize == 56
    assert C.sizeof(A.HitDesc) == 88 and A.hit_dtype().itemsize == 88
    assert C.sizeof(A.JpegInfo) == 16 * 4 + 3 * 64 * 2 + 4 * 8          # rt_jpeg_info


def test_error_codes_and_messages(rt):
    A = rt._abi
    lib = A.lib()
    s = rt.Scene()
    assert lib.rt_mat_lambertian(s._b, 99) == A.RT_ERR_OUT_OF_RANGE
    assert b"unknown texture" in lib.rt_last_error()
    assert lib.rt_hit_sphere(s._b, (C.c_double * 3)(0, 0, 0), 1.0, 5) == A.RT_ERR_OUT_OF_RANGE
    assert lib.rt_hit_bvh(s._b, (C.c_int * 1)(0), 0) == A.RT_ERR_INVALID_ARGUMENT
    assert lib.rt_tex_solid(None, 0.0, 0.0, 0.0) == A.RT_ERR_INVALID_ARGUMENT
    with pytest.raises(A.RtError):
        s.Lambertian(rt.Handle(42))
    cs = rt.CameraSettings(image_width=0)
    with pytest.raises(A.RtError):
        rt.Camera(cs)


def test_product_never_touches_the_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "rust-tracing_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"\boracle\b", txt) and "oracle" in re.sub(r"(#|//).*", "", txt):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders
    for f in os.listdir(os.path.join(ROOT, "include")):
        assert "liboracle" not in open(os.path.join(ROOT, "include", f)).read()


def test_product_library_reads_no_environment():
    """The drop-in's behaviour must not depend on the host process's environment: every getenv in the library's sources
    sits inside an #ifdef RT_B200_DEV block (the A/B build), and the default build does not define it."""
    csrc = os.path.join(ROOT, "rust-tracing_b200", "csrc")
    for dirpath, _, files in os.walk(csrc):
        for f in files:
            if not f.endswith((".cpp", ".cu", ".cuh", ".h")):
                continue
            depth = 0
            for line in open(os.path.join(dirpath, f), errors="replace"):
                t = line.strip()
                if t.startswith("#ifdef RT_B200_DEV"):
                    depth += 1
                elif depth and t.startswith(("#ifdef", "#ifndef", "#if ")):
                    depth += 1
                elif depth and t.startswith("#endif"):
                    depth -= 1
                elif "getenv" in t and not t.startswith("//"):
                    assert depth > 0, f"{f}: getenv outside RT_B200_DEV: {t}"
    assert "RT_B200_DEV" not in open(os.path.join(ROOT, "rust-tracing_b200", "build.py")).read().split("def build(")[0]


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="GPU present")
def test_no_cpu_fallback_without_gpu(rt):
    with pytest.raises(rt._abi.RtError) as e:
        rt.Context(0)
    assert e.value.status == rt._abi.RT_ERR_NO_DEVICE


def test_cpp_mirror_header_builds_and_runs(rt, tmp_path):
    """include/rt_b200.hpp (the C++ mirror of the crate's constructors) compiles standalone against the C ABI and the
    example scene flattens; without a GPU the program stops at rt_context_create (no CPU fallback)."""
    import subprocess
    exe = str(tmp_path / "cornell")
    csrc = os.path.join(ROOT, "rust-tracing_b200", "csrc")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "cornell.cpp"), "-L" + csrc, "-lrt_b200", "-Wl,-rpath," + csrc, "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe, "2"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "25 hittables, 15 bvh nodes -> 48 stream words (2 inner, 6 quad, 2 box, 2 instance ops)" in out.stdout
    assert ("no GPU" in out.stdout) or ("rendered 200x200" in out.stdout)
