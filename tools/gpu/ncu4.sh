set -x
mkdir -p gpurun_out
timeout 300 python tools/ncu_target.py --spp 32 --reps 2 > gpurun_out/r2_ncu_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/prof_r2_f -f python tools/ncu_target.py --spp 32 --reps 2 > gpurun_out/r2_ncu_f.log 2>&1
tail -3 gpurun_out/r2_ncu_f.log
