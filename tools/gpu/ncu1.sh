set -x
mkdir -p gpurun_out
python tools/ncu_target.py --spp 32 --reps 2 > gpurun_out/r2_ncu_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/prof_r2_a -f python tools/ncu_target.py --spp 32 --reps 2 > gpurun_out/r2_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/prof_r2_a_cornell -f python tools/ncu_target.py --scene 6 --spp 64 --reps 2 > gpurun_out/r2_ncu_a_cornell.log 2>&1
ls -la gpurun_out
