mkdir -p gpurun_out
L=rust-tracing_b200/csrc
rm -f gpurun_out/r2_ab23.log
for s in 8 0 3; do timeout 400 python tools/ab_lib.py --scene $s --spp 300 --rounds 3 $L/librt_b200.so $L/librt_b200_sm3.so $L/librt_b200_sm6.so 2>&1 | tail -4 >> gpurun_out/r2_ab23.log; done
cat gpurun_out/r2_ab23.log
