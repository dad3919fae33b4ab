mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -q -x > gpurun_out/r2_pytest6.log 2>&1; echo "rc $?" >> gpurun_out/r2_pytest6.log; tail -15 gpurun_out/r2_pytest6.log
timeout 300 python tools/sweep_dev.py --scene 8 --spp 200 --rounds 2 RT_B200_KERNEL=mk RT_B200_XCHG_MIN=4 RT_B200_XCHG_MIN=12 RT_B200_SLAB_DROP=4 RT_B200_SLAB_DROP=12 RT_B200_MIN_TRAV=16 RT_B200_MIN_TRAV=28 RT_B200_SLAB_FAST=10 > gpurun_out/r2_q1.log 2>&1; echo "rc $?" >> gpurun_out/r2_q1.log; cat gpurun_out/r2_q1.log
