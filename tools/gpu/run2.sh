mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1; tail -15 gpurun_out/r2_pytest3.log
timeout 900 python bench.py > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 3000 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
