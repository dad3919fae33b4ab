mkdir -p gpurun_out
python tools/sweep_dev.py --scene 8 --spp 400 RT_B200_SHADE_MIN=16 RT_B200_SHADE_MIN=20 RT_B200_SHADE_MIN=28 RT_B200_SHADE_MIN=32 RT_B200_SLAB_FAST=10 RT_B200_SLAB_FAST=12 RT_B200_SLAB_FAST=16 RT_B200_SLAB_FAST=20 RT_B200_SPHERE_REPS=1 RT_B200_SPHERE_REPS=4 RT_B200_OPS_GLOBAL=1 RT_B200_CHUNK=16 RT_B200_CHUNK=64 > gpurun_out/r2_sweep1.log 2>&1
cat gpurun_out/r2_sweep1.log
python tools/sweep_dev.py --scene 6 --spp 400 RT_B200_SHADE_MIN=16 RT_B200_SHADE_MIN=28 RT_B200_SLAB_FAST=10 RT_B200_SLAB_FAST=18 RT_B200_OPS_GLOBAL=1 > gpurun_out/r2_sweep1_cornell.log 2>&1
cat gpurun_out/r2_sweep1_cornell.log
python tools/sweep_dev.py --scene 0 --spp 400 RT_B200_SHADE_MIN=16 RT_B200_SHADE_MIN=28 RT_B200_SLAB_FAST=10 RT_B200_SLAB_FAST=18 RT_B200_OPS_GLOBAL=1 > gpurun_out/r2_sweep1_balls.log 2>&1
cat gpurun_out/r2_sweep1_balls.log
