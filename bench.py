#!/usr/bin/env python
"""bench.py — Mpaths/s (pixels x spp / s) of the ray_color hot path on the Book-2 final_scene,
800x800, 10000 spp, depth 40 (BASELINE.json configs[4], the config the metric is quoted on).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the CPU reference arm (oracle port)

A "step" is one full render of the workload: every rank accumulates its share of the 10000 samples of every
pixel into a device float4 SUM framebuffer (rt_render_accumulate, scene already resident in HBM) and the partial
framebuffers are summed onto rank 0 with one NCCL reduce. `value` is timed with CUDA events on the launching
stream, max over ranks. `e2e` is the same workload through the C ABI with host buffers: rt_scene_upload (host
scene description incl. the 61 MB RGB8 earth image -> device) + render + the SUM framebuffer back in host memory.
One JSON line is printed by rank 0; at N = 1 it also carries `configs`: every other BASELINE.json config measured in
the same process after the headline (each a fraction of a second on the GPU), with its own roofline fraction.

At N > 1 the sample range is split in proportion to each rank's warm-up throughput (GPUs of one box differ by a few
per cent on this kernel; the slowest rank sets the step time): rust_tracing_b200.distributed.shard_samples_weighted.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mpaths/sec (pixels x spp / s) on Book-2 final_scene"
UNIT = "Mpaths/s"
WORKLOAD = {"scene": 8, "name": "final_scene", "width": 800, "height": 800, "spp": 10000, "max_depth": 40}
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # SMs x lanes x FMA x max SM clock
# BASELINE.json configs[0..3] (+ the three textured scenes of configs[1]); name: (scene, width, spp, depth override)
OTHER_CONFIGS = {"cfg1_random_balls": (0, 400, 100, 50), "cfg2a_checker": (1, 800, 500, 0), "cfg2b_earth": (2, 800, 500, 0),
                 "cfg2c_perlin": (3, 800, 500, 0), "cfg3_cornell_box": (6, 600, 1000, 50), "cfg4_cornell_smoke": (7, 600, 2000, 0)}


def flops_per_path():
    p = os.path.join(ROOT, "profiles", "flops_per_path.json")
    try:
        with open(p) as f:
            return float(json.load(f)["configs"]["cfg5_final_scene"]["flops_per_path"]), "profiles/flops_per_path.json"
    except Exception:
        return None, "missing"


# SURVEY.md §8(d) cost table applied to the ops the DEVICE traversal executes (rt_render_count_ops), so that
# roofline.achieved counts arithmetic the kernel really does. The reference visits far more nodes for the same
# image (per-axis AABB quirk, origin-inclusive list boxes, 6 quads per box), see profiles/flops_per_path.json.
DEVICE_COST = {"slab": 24, "box": 24, "box_hit": 8, "sphere": 24, "sphere_moving": 6, "sphere_hit": 8, "quad": 16, "quad_hit": 50,
               "xform_enter": 15, "finalize_xform": 12, "medium": 56, "medium_hit": 8, "lambertian": 42 + 30, "metal": 57 + 30,
               "dielectric": 60 + 30, "isotropic": 33 + 30, "light": 30, "tex_noise": 850, "tex_image": 12, "tex_checker": 9,
               "segments": 9 + 3, "paths": 31}


def device_flops_per_path(counts):
    return sum(DEVICE_COST.get(k, 0) * v for k, v in counts.items()) / max(1, counts["paths"])


def ncu_traffic_per_launch():
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1])); mx.append(float(parts[2])); power.append(float(parts[3]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


def build_scene(rt, spp):
    earth, earth_src = rt.load_earth()
    s, cs = rt.builtin_scene(WORKLOAD["scene"], image_width=WORKLOAD["width"], samples_per_pixel=spp,
                             max_depth=WORKLOAD["max_depth"], scene_seed=1, bvh_seed=2, perlin_seed=3, earth=earth)
    cam = rt.Camera(cs)
    assert cam.shape == (WORKLOAD["height"], WORKLOAD["width"])
    return s, cam, earth, earth_src


def scene_h2d_bytes(rt, s, cam):
    import ctypes as C
    A = rt._abi
    d = s.desc
    n = (d.n_textures * C.sizeof(A.TextureDesc) + d.n_materials * C.sizeof(A.MaterialDesc)
         + d.n_hittables * C.sizeof(A.HittableDesc) + d.n_list_items * 4 + d.n_bvh_nodes * C.sizeof(A.BvhNodeDesc)
         + d.n_perlins * C.sizeof(A.PerlinDesc) + C.sizeof(A.CameraDesc))
    for k in range(d.n_images):
        n += d.images[k].width * d.images[k].height * 3
    return n


def cpu_reference_step(ob, s, cam, spp, threads=0):
    t0 = time.perf_counter()
    _, cnt = ob.render(s.desc, cam, 0, spp, seed=0, mode=0, threads=threads)
    dt = time.perf_counter() - t0
    return cnt["paths"], dt


def pick_cpu_spp(ob, s, cam, target_s):
    """Bounded CPU sample: time 1 spp of the full 800x800 frame, then pick the spp that makes ~target_s."""
    paths, dt = cpu_reference_step(ob, s, cam, 1)
    spp = max(1, min(64, int(round(target_s / max(dt, 1e-3)))))
    return spp, paths / dt / 1e6


def run_reference(args):
    """The reference arm: the reference's own algorithm on the host cores. The Rust crate cannot be built in this
    image (no cargo/rustc; DESIGN.md "Oracle"), so this is the C++ f64 restatement (oracle/, kind "port") with
    all host threads, each step a bounded sample (same scene/size/depth, reduced spp; Mpaths/s does not depend on spp)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import rust_tracing_b200 as rt
    from oracle import binding as ob
    s, cam, _, earth_src = build_scene(rt, WORKLOAD["spp"])
    cores = os.cpu_count() or 1
    spp, _ = pick_cpu_spp(ob, s, cam, 8.0)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_step(ob, s, cam, 1)
    total_paths, total_t = 0, 0.0
    for _ in range(args.steps):
        p, dt = cpu_reference_step(ob, s, cam, spp)
        total_paths += p
        total_t += dt
    value = total_paths / total_t / 1e6
    sample = f"final_scene 800x800 depth 40 at {spp} spp per step ({args.steps} steps), f64, {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_t / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": f"synthetic scene (seeded layout); earth texels: {earth_src}",
            # the same workload as the device arm names; what a step of THIS arm covers of it is the bounded sample
            "config": {"workload": f"Book-2 final_scene {WORKLOAD['width']}x{WORKLOAD['height']}, {WORKLOAD['spp']} spp, depth {WORKLOAD['max_depth']} (BASELINE.json configs[4])",
                       "sample": f"bounded: {spp} of the {WORKLOAD['spp']} spp per step (Mpaths/s does not depend on spp)", "spp_per_step": spp},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


def measure_other_configs(rt, ctx, torch, earth, peak_tflops):
    """BASELINE.json's other configs at their full size and spp on this GPU: CUDA events around rt_render_accumulate (scene
    resident, 2 warm-up launches, best of 3, the L2 flushed in between), device-counted flops per path, roofline fraction."""
    dev = torch.device(f"cuda:{ctx.device_id}")
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    for name, (idx, width, spp, depth) in OTHER_CONFIGS.items():
        s, cs = rt.builtin_scene(idx, image_width=width, samples_per_pixel=spp, max_depth=depth, earth=earth if idx in (2, 8) else None)
        cam = rt.Camera(cs)
        h, w = cam.shape
        ds = ctx.upload(s)
        fb = torch.zeros((h, w, 4), dtype=torch.float32, device=dev)
        for _ in range(2):
            ctx.render_accumulate(ds, cam, 0, spp, 0, fb.data_ptr(), stream.cuda_stream)
        best = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(3):
            flush.zero_()
            fb.zero_()
            torch.cuda.synchronize()
            e0.record(stream)
            ctx.render_accumulate(ds, cam, 0, spp, 1 + rep, fb.data_ptr(), stream.cuda_stream)
            e1.record(stream)
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        ok = bool((fb[..., 3] == spp).all().item())
        ops = ctx.count_ops(ds, cam, 0, 4, seed=0)
        fpp = device_flops_per_path(ops)
        mpaths = h * w * spp / best / 1e3
        out[name] = {"scene": rt.SCENE_NAMES[idx], "width": w, "height": h, "spp": spp, "max_depth": int(cam.max_depth),
                     "mpaths": mpaths, "ms": best, "flops_per_path": fpp, "tflops": fpp * mpaths * 1e6 / 1e12,
                     "roofline_frac": fpp * mpaths * 1e6 / 1e12 / peak_tflops, "sample_count_ok": ok}
        ds.close()
    return out


_REAL_STDOUT = None


def claim_stdout():
    """Everything any library prints on fd 1 (NCCL's version banner, ...) goes to stderr; the ONE JSON line is written
    to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--spp", type=int, default=WORKLOAD["spp"], help="debug only: anything but 10000 marks the line invalid")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config rows (BASELINE.json configs[0..3]) after the headline")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import rust_tracing_b200 as rt
    from rust_tracing_b200.distributed import shard_samples, shard_samples_weighted, reduce_to_root

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    spp = args.spp
    s, cam, earth, earth_src = build_scene(rt, spp)
    ctx = rt.Context(local)
    ds = ctx.upload(s)
    h, w = cam.shape
    fb = torch.zeros((h, w, 4), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    begin, count = shard_samples(spp, rank, world)
    seed = 0

    def step(timed):
        """One full render: zero the SUM buffer, accumulate this rank's sample share, reduce to rank 0."""
        flush.zero_()                                   # L2 flush between iterations (outside the timed events)
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        fb.zero_()
        if count > 0:
            ctx.render_accumulate(ds, cam, begin, count, seed, fb.data_ptr(), stream.cuda_stream)
        e1.record(stream)
        reduce_to_root(fb)
        e2.record(stream)
        e2.synchronize()
        return e0.elapsed_time(e2), e0.elapsed_time(e1)

    share_note = "equal"
    for k in range(args.warmup):
        _, kms = step(False)
        if world > 1 and k == args.warmup - 1:
            # the last warm-up step doubles as the calibration: split the samples in proportion to each rank's speed
            rate = torch.zeros(world, dtype=torch.float64, device=dev)
            rate[rank] = count / max(kms, 1e-6)
            dist.all_reduce(rate)
            weights = [float(x) for x in rate.cpu()]
            begin, count = shard_samples_weighted(spp, weights)[rank]
            share_note = "proportional to warm-up throughput: " + "/".join(str(c) for _, c in shard_samples_weighted(spp, weights))
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    step_ms, kernel_ms = [], []
    for _ in range(args.steps):
        barrier()
        a, b = step(True)
        step_ms.append(a)
        kernel_ms.append(b)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_total = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    t_kernel = torch.tensor([sum(kernel_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_total, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_kernel, op=dist.ReduceOp.MAX)
    total_ms = float(t_total.item())
    paths_per_step = h * w * spp
    value = paths_per_step * args.steps / (total_ms * 1e-3) / 1e6
    check = None
    if rank == 0:
        res = fb.cpu().numpy()
        ok_count = bool(np.all(res[..., 3] == spp))
        check = {"sample_count_ok": ok_count, "mean_rgb": [float(x) for x in res[..., :3].mean(axis=(0, 1)) / spp]}

    # ---- e2e: through the C ABI with host buffers, every step: scene upload + render + framebuffer to host
    pinned = torch.empty((h, w, 4), dtype=torch.float32).pin_memory()
    h2d = scene_h2d_bytes(rt, s, cam)

    def e2e_step():
        ds2 = ctx.upload(s)                             # H2D: flat scene description + RGB8 earth image
        if world == 1:                                  # the literal drop-in call: rt_render(ctx, scene, cam, ..., host buffer)
            host = ctx.render(ds2, cam, begin, count, seed)
            ds2.close()
            return host
        fb.zero_()
        if count > 0:
            ctx.render_accumulate(ds2, cam, begin, count, seed, fb.data_ptr(), stream.cuda_stream)
        reduce_to_root(fb)
        if rank == 0:
            pinned.copy_(fb, non_blocking=False)        # D2H: the SUM framebuffer (renderer.rs:49 `collect`)
        torch.cuda.synchronize()
        ds2.close()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = paths_per_step * args.steps / float(e2e_t.item()) / 1e6

    if rank == 0:
        fpp, fpp_src = flops_per_path()
        try:
            peak = ctx.measure_fp32_peak()
            peak_src = "measured live: FFMA loop kernel (rt_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 figure"
        except Exception:
            peak, peak_src = NOMINAL_FP32_TFLOPS, "nominal 148 SMs x 128 x 2 x 1.965 GHz"
        kern_s = float(t_kernel.item()) * 1e-3 / args.steps
        paths_per_launch = h * w * count
        ops = ctx.count_ops(ds, cam, 0, 4, seed=0)          # instrumented kernel, outside every timed region
        dfpp = device_flops_per_path(ops)
        roofline = {"bound": "fp32", "kernel": "render_kernel_mk", "unit": "TFLOP/s", "peak": peak, "peak_source": peak_src,
                    "peak_nominal": NOMINAL_FP32_TFLOPS,
                    "achieved": dfpp * paths_per_launch / kern_s / 1e12, "frac": dfpp * paths_per_launch / kern_s / 1e12 / peak,
                    "traffic": None, "flops_per_path": dfpp,
                    "flops_per_path_source": "device op counts (rt_render_count_ops, 4 spp) x SURVEY.md 8(d) cost table",
                    "device_ops_per_path": {k: v / ops["paths"] for k, v in ops.items()},
                    "kernel_ms_per_launch": kern_s * 1e3}
        if fpp:   # the same image at the reference's own visit counts (oracle-counted): a work-rate, not a utilisation
            roofline["reference_flops_per_path"] = fpp
            roofline["reference_flops_per_path_source"] = fpp_src
            roofline["reference_equivalent_tflops"] = fpp * paths_per_launch / kern_s / 1e12
        tr = ncu_traffic_per_launch()
        if tr:   # DRAM bytes per path from the committed ncu --set full capture, scaled to this launch's path count
            roofline["traffic"] = tr.get("dram_bytes_per_path", 0.0) * paths_per_launch
            roofline["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per path from profiles/ncu_traffic.json ("
                                        + str(tr.get("capture")) + ") x paths of this launch; HBM is not the bound of this kernel")
        configs = None
        if world == 1 and not args.no_configs:
            configs = measure_other_configs(rt, ctx, torch, earth, peak)
            configs["cfg5_final_scene"] = {"scene": "final_scene", "width": w, "height": h, "spp": spp, "mpaths": value,
                                           "ms": total_ms / args.steps, "flops_per_path": dfpp, "tflops": roofline["achieved"],
                                           "roofline_frac": roofline["frac"]}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import binding as ob
            cores = os.cpu_count() or 1
            cspp, _ = pick_cpu_spp(ob, s, cam, 15.0)
            p, dt = cpu_reference_step(ob, s, cam, cspp)
            cpu = {"value": p / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"final_scene 800x800 depth 40 at {cspp} spp ({p} paths, {dt:.1f} s), C++ f64 restatement of rust-tracing, {cores} threads"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": f"synthetic scene (seeded layout, BASELINE.md seeds); earth texels: {earth_src}",
                "config": {"workload": f"Book-2 final_scene {w}x{h}, {spp} spp, depth {WORKLOAD['max_depth']} (BASELINE.json configs[4])",
                           "parallelism": f"sample-range sharding x{world} ({share_note}), one NCCL reduce to rank 0",
                           "l2": "flushed between timed iterations (256 MiB memset)", "valid": spp == WORKLOAD["spp"]},
                "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h * w * 16},
                "gpu_launches": args.steps * (1 if count > 0 else 0),   # render_kernel_mk, once per step on this rank
                "check": check}
        if configs:
            line["configs"] = configs
        emit(line)
    barrier()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
