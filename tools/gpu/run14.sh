mkdir -p gpurun_out
L=rust-tracing_b200/csrc
timeout 400 python tools/ab_lib.py --scene 8 --spp 400 --rounds 3 $L/librt_b200.so $L/librt_b200_t640.so $L/librt_b200_t896.so $L/librt_b200_t1024.so > gpurun_out/r2_ab9_threads.log 2>&1; tail -5 gpurun_out/r2_ab9_threads.log
timeout 600 python tools/sweep_dev.py --scene 8 --spp 300 --rounds 3 RT_B200_KERNEL=q RT_B200_SHADE_MIN=20 RT_B200_SHADE_MIN=28 RT_B200_SLAB_FAST=4 RT_B200_SLAB_FAST=8 RT_B200_SLAB_FAST=10 RT_B200_SPHERE_REPS=1 RT_B200_SPHERE_REPS=3 RT_B200_CHUNK=4 RT_B200_CHUNK=16 RT_B200_OPS_GLOBAL=1 > gpurun_out/r2_sweep5_after_icache.log 2>&1; cat gpurun_out/r2_sweep5_after_icache.log
