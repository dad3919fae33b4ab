mkdir -p gpurun_out
L=rust-tracing_b200/csrc
timeout 600 python -m pytest tests/test_gpu_jpeg.py tests/test_gpu_bvh_build.py -m gpu -q -x > gpurun_out/r2_pytest10.log 2>&1; echo "rc $?" >> gpurun_out/r2_pytest10.log; tail -5 gpurun_out/r2_pytest10.log
timeout 300 python tools/ncu_small_kernels.py > gpurun_out/r2_small_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none -k regex:'expand_image|finalize|jpeg|bvh' -o gpurun_out/prof_r2_small -f python tools/ncu_small_kernels.py > gpurun_out/r2_small_ncu.log 2>&1
cat gpurun_out/r2_small_plain.log
timeout 400 python tools/ab_lib.py --scene 8 --spp 400 --rounds 3 $L/librt_b200_base.so $L/librt_b200_vote2.so > gpurun_out/r2_ab6_vote_ballots.log 2>&1; tail -3 gpurun_out/r2_ab6_vote_ballots.log
for s in 0 6; do timeout 200 python tools/ab_lib.py --scene $s --spp 400 --rounds 2 $L/librt_b200_base.so $L/librt_b200_vote2.so 2>&1 | tail -2 >> gpurun_out/r2_ab6_scenes.log; done; cat gpurun_out/r2_ab6_scenes.log
