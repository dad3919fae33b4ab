mkdir -p gpurun_out
for s in 6 7 4 8 0; do python tools/sweep_dev.py --scene $s --spp 400 RT_B200_QUAD_REPS=1 RT_B200_QUAD_REPS=2 RT_B200_QUAD_REPS=8 RT_B200_SLAB_FAST=10,RT_B200_CHUNK=16 RT_B200_SLAB_FAST=8 >> gpurun_out/r2_sweep2.log 2>&1; done
cat gpurun_out/r2_sweep2.log
