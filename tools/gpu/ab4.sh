mkdir -p gpurun_out
L=rust-tracing_b200/csrc
for s in 8 0 6 7; do python tools/ab_lib.py --scene $s --spp 400 --rounds 3 $L/librt_b200_m2.so $L/librt_b200.so $L/librt_b200_m5.so 2>&1 | tail -3 >> gpurun_out/r2_ab4.log; done
cat gpurun_out/r2_ab4.log
