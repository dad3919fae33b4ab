mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -q -x > gpurun_out/r2_pytest7.log 2>&1; echo "rc $?" >> gpurun_out/r2_pytest7.log; tail -15 gpurun_out/r2_pytest7.log
timeout 300 python tools/sweep_dev.py --scene 8 --spp 200 --rounds 2 RT_B200_KERNEL=mk RT_B200_SLAB_DROP=8 RT_B200_SLAB_DROP=12 RT_B200_SLAB_DROP=20 RT_B200_SLAB_DROP=24 RT_B200_MIN_TRAV=6 RT_B200_MIN_TRAV=20 RT_B200_SPHERE_REPS=1 RT_B200_SPHERE_REPS=4 > gpurun_out/r2_q2.log 2>&1; echo "rc $?" >> gpurun_out/r2_q2.log; cat gpurun_out/r2_q2.log
