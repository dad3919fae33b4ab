set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_final.log 2>&1; echo "rc $?" >> gpurun_out/r2_pytest_final.log; tail -4 gpurun_out/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "ref rc $?"
timeout 300 python tools/ncu_target.py --spp 32 --reps 2 > gpurun_out/r2_ncu_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -o gpurun_out/prof_r2_h -f python tools/ncu_target.py --spp 32 --reps 2 > gpurun_out/r2_ncu_h.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.json 2> gpurun_out/r2_bench_under_ncu.err; echo "launch list rc $?"
head -c 1200 gpurun_out/r2_bench_1gpu.json
