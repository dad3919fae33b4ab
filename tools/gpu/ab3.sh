mkdir -p gpurun_out
L=rust-tracing_b200/csrc
python tools/ab_lib.py --scene 8 --spp 500 --rounds 3 $L/librt_b200_r1.so $L/librt_b200_m2.so $L/librt_b200_m4.so > gpurun_out/r2_ab3.log 2>&1
tail -4 gpurun_out/r2_ab3.log
for s in 0 6 7; do python tools/ab_lib.py --scene $s --spp 500 --rounds 2 $L/librt_b200_r1.so $L/librt_b200_m2.so $L/librt_b200_m4.so 2>&1 | tail -3 >> gpurun_out/r2_ab3_scenes.log; done
cat gpurun_out/r2_ab3_scenes.log
