#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench rc $?"; cut -c1-120 gpurun_out/bench_1gpu.json
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_for_ncu.json 2> gpurun_out/bench_for_ncu.err && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err; echo "launch list rc $?"
timeout 120 python tools/ncu_target.py --spp 32 --reps 2 > gpurun_out/ncu_plain.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/prof_r1_g -f python tools/ncu_target.py --spp 32 --reps 2 > gpurun_out/ncu_g.log 2>&1; echo "ncu full rc $?"
for k in 1 2; do
  echo "== default (fold per scene)"; timeout 100 python tools/ncu_target.py --spp 1000 --reps 2 | grep rep
  echo "== RT_B200_FOLD_BOX=0"; RT_B200_FOLD_BOX=0 timeout 100 python tools/ncu_target.py --spp 1000 --reps 2 | grep rep
done
