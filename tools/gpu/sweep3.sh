mkdir -p gpurun_out
for s in 8 0 6; do python tools/sweep_dev.py --scene $s --spp 400 RT_B200_SLAB_FAST=4 RT_B200_SLAB_FAST=6 RT_B200_SLAB_FAST=12 RT_B200_CHUNK=8 RT_B200_CHUNK=32 RT_B200_SHADE_MIN=22 RT_B200_SHADE_MIN=26 RT_B200_SPHERE_REPS=3 >> gpurun_out/r2_sweep3.log 2>&1; done
cat gpurun_out/r2_sweep3.log
