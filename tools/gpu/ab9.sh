mkdir -p gpurun_out
L=rust-tracing_b200/csrc
rm -f gpurun_out/r2_ab19.log
for s in 8 0 6; do timeout 400 python tools/ab_lib.py --scene $s --spp 400 --rounds 3 $L/librt_b200.so $L/librt_b200_rag.so $L/librt_b200_spec.so $L/librt_b200_both.so 2>&1 | tail -5 >> gpurun_out/r2_ab19.log; done
cat gpurun_out/r2_ab19.log
