mkdir -p gpurun_out
L=rust-tracing_b200/csrc
rm -f gpurun_out/r2_ab20.log
for s in 8 6 7 0; do timeout 400 python tools/ab_lib.py --scene $s --spp 400 --rounds 3 $L/librt_b200.so $L/librt_b200_park.so 2>&1 | tail -3 >> gpurun_out/r2_ab20.log; done
cat gpurun_out/r2_ab20.log
RT_B200_LIB=$L/librt_b200_park.so timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
