mkdir -p gpurun_out
rm -f gpurun_out/r2_sweep7.log
for s in 8 0 6 7; do timeout 300 python tools/sweep_dev.py --scene $s --spp 300 --rounds 3 RT_B200_SPHERE_REPS=3 RT_B200_SPHERE_REPS=4 RT_B200_QUAD_REPS=12 RT_B200_QUAD_REPS=16 RT_B200_SHADE_MIN=22 RT_B200_SHADE_MIN=26 RT_B200_SLAB_FAST=5 RT_B200_SLAB_FAST=7 >> gpurun_out/r2_sweep7.log 2>&1; done
cat gpurun_out/r2_sweep7.log
