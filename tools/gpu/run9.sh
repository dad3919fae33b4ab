mkdir -p gpurun_out
L=rust-tracing_b200/csrc
timeout 600 python -m pytest tests/test_gpu_render.py tests/test_gpu_hits.py -m gpu -q -x > gpurun_out/r2_pytest9.log 2>&1; echo "rc $?" >> gpurun_out/r2_pytest9.log; tail -3 gpurun_out/r2_pytest9.log
timeout 300 python tools/ab_lib.py --scene 8 --spp 400 --rounds 3 $L/librt_b200_head.so $L/librt_b200_dev.so --env RT_B200_KERNEL=mk > gpurun_out/r2_ab5_box_reject_fold.log 2>&1; tail -3 gpurun_out/r2_ab5_box_reject_fold.log
for s in 0 6 7; do timeout 200 python tools/ab_lib.py --scene $s --spp 400 --rounds 2 $L/librt_b200_head.so $L/librt_b200_dev.so --env RT_B200_KERNEL=mk 2>&1 | tail -2 >> gpurun_out/r2_ab5_scenes.log; done; cat gpurun_out/r2_ab5_scenes.log
timeout 300 python tools/sweep_dev.py --scene 8 --spp 200 --rounds 2 RT_B200_KERNEL=mk RT_B200_SLAB_DROP=12 RT_B200_SLAB_DROP=16 RT_B200_MIN_TRAV=16 RT_B200_MIN_TRAV=20 > gpurun_out/r2_q4.log 2>&1; echo "rc $?" >> gpurun_out/r2_q4.log; cat gpurun_out/r2_q4.log
RT_B200_LIB=rust-tracing_b200/csrc/librt_b200_dev768.so timeout 300 python tools/sweep_dev.py --scene 8 --spp 200 --rounds 2 RT_B200_KERNEL=mk RT_B200_SLAB_DROP=12 RT_B200_MIN_TRAV=16 RT_B200_MIN_TRAV=20 > gpurun_out/r2_q4_768.log 2>&1; echo "rc $?" >> gpurun_out/r2_q4_768.log; cat gpurun_out/r2_q4_768.log
