mkdir -p gpurun_out
L=rust-tracing_b200/csrc
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest13.log 2>&1; echo "rc $?" >> gpurun_out/r2_pytest13.log; tail -5 gpurun_out/r2_pytest13.log
timeout 400 python tools/ab_lib.py --scene 8 --spp 400 --rounds 3 $L/librt_b200_nofd.so $L/librt_b200_fd.so $L/librt_b200.so > gpurun_out/r2_ab8_small_code.log 2>&1; tail -4 gpurun_out/r2_ab8_small_code.log
for s in 0 3 6 7; do timeout 200 python tools/ab_lib.py --scene $s --spp 400 --rounds 2 $L/librt_b200_nofd.so $L/librt_b200_fd.so $L/librt_b200.so 2>&1 | tail -3 >> gpurun_out/r2_ab8_scenes.log; done; cat gpurun_out/r2_ab8_scenes.log
