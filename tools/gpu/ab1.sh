set -x
mkdir -p gpurun_out
L=rust-tracing_b200/csrc
python tools/ab_lib.py --scene 8 --spp 500 --rounds 3 $L/librt_b200_r1.so $L/librt_b200.so $L/librt_b200_t512.so $L/librt_b200_t640.so $L/librt_b200_t896.so > gpurun_out/r2_ab1.log 2>&1
cat gpurun_out/r2_ab1.log
for s in 0 6 7 3; do python tools/ab_lib.py --scene $s --spp 500 --rounds 2 $L/librt_b200_r1.so $L/librt_b200.so >> gpurun_out/r2_ab1_scenes.log 2>&1; done
cat gpurun_out/r2_ab1_scenes.log
