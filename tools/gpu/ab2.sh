mkdir -p gpurun_out
L=rust-tracing_b200/csrc
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; tail -3 gpurun_out/r2_pytest2.log
python tools/ab_lib.py --scene 8 --spp 500 --rounds 3 $L/librt_b200_r1.so $L/librt_b200.so $L/librt_b200_t640.so > gpurun_out/r2_ab2.log 2>&1
cat gpurun_out/r2_ab2.log | tail -4
for s in 0 6 7; do python tools/ab_lib.py --scene $s --spp 500 --rounds 2 $L/librt_b200_r1.so $L/librt_b200.so 2>&1 | tail -2 >> gpurun_out/r2_ab2_scenes.log; done
cat gpurun_out/r2_ab2_scenes.log
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/prof_r2_b -f python tools/ncu_target.py --spp 32 --reps 2 > gpurun_out/r2_ncu_b.log 2>&1
