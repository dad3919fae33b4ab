mkdir -p gpurun_out
L=rust-tracing_b200/csrc
rm -f gpurun_out/r2_ab21.log
for s in 6 7 4 8; do timeout 400 python tools/ab_lib.py --scene $s --spp 400 --rounds 3 $L/librt_b200.so $L/librt_b200_hq.so 2>&1 | tail -3 >> gpurun_out/r2_ab21.log; done
cat gpurun_out/r2_ab21.log
