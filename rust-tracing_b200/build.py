"""Builds csrc/librt_b200.so: host scene code (g++ via nvcc) + sm_100a kernels, in-tree.

    python rust-tracing_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU. The .so is git-ignored but travels to the GPU box with gpurun.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "librt_b200.so")
SOURCES = [
    os.path.join(CSRC, "host", "scene_builder.cpp"),
    os.path.join(CSRC, "host", "scenes.cpp"),
    os.path.join(CSRC, "host", "jpeg_entropy.cpp"),
    os.path.join(CSRC, "device", "scene_compile.cpp"),
    os.path.join(CSRC, "device", "rt_cuda.cu"),
]
DEPS = SOURCES + [
    os.path.join(CSRC, "host", "host_common.h"),
    os.path.join(CSRC, "host", "host_rng.h"),
    os.path.join(CSRC, "device", "dev_scene.h"),
    os.path.join(CSRC, "device", "rt_kernels.cuh"),
    os.path.join(CSRC, "device", "render_mk.cuh"),
    os.path.join(CSRC, "device", "render_q.cuh"),
    os.path.join(CSRC, "device", "bvh_build.cuh"),
    os.path.join(CSRC, "device", "jpeg_kernels.cuh"),
    os.path.join(CSRC, "host", "jpeg_entropy.h"),
    os.path.join(HERE, "..", "include", "rt_b200.h"),
    os.path.join(HERE, "..", "include", "rt_b200.hpp"),
]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -prec-div=false -prec-sqrt=false: device divisions / square roots use the 1-2 ulp fast sequences instead of the
# IEEE-rounded ones (+8-10% Mpaths/s on every scene, profiles/r1_ab12*.log); the GPU parity suite passes unchanged
# within its stated tolerances. Denormals and FMA contraction keep their defaults.
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-prec-div=false", "-prec-sqrt=false",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-shared", "-Xptxas", "-v"]


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force=False, verbose=False, extra_flags=(), out=None):
    """extra_flags / out: development variants (e.g. -DRT_OPT_... into csrc/librt_b200_<name>.so, selected at run time
    with RT_B200_LIB=<path>); the product is always the default build."""
    if out is None and not force and up_to_date():
        return OUT
    cmd = [NVCC] + FLAGS + list(extra_flags) + ["-o", out or OUT] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(CSRC, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed (see {log})")
    return out or OUT


if __name__ == "__main__":
    if "--variant" in sys.argv:      # python build.py --variant NAME -DFLAG ...   -> csrc/librt_b200_NAME.so
        k = sys.argv.index("--variant")
        name = sys.argv[k + 1]
        print(build(force=True, verbose="-v" in sys.argv, extra_flags=[a for a in sys.argv[k + 2:] if a != "-v"],
                    out=os.path.join(CSRC, f"librt_b200_{name}.so")))
    else:
        build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
        print(OUT)
