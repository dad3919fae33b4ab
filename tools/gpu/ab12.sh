mkdir -p gpurun_out
L=rust-tracing_b200/csrc
rm -f gpurun_out/r2_ab22.log
for s in 8 0 3 7; do timeout 400 python tools/ab_lib.py --scene $s --spp 400 --rounds 3 $L/librt_b200.so $L/librt_b200_hs.so 2>&1 | tail -3 >> gpurun_out/r2_ab22.log; done
cat gpurun_out/r2_ab22.log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
