set -x
mkdir -p gpurun_out
timeout 900 python tools/fuzz_render.py --seeds 64 > gpurun_out/r2_fuzz_render.log 2>&1; echo "rc $?" >> gpurun_out/r2_fuzz_render.log; grep -v "ok " gpurun_out/r2_fuzz_render.log | tail -12
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_after_fix.log 2>&1; echo "rc $?" >> gpurun_out/r2_pytest_after_fix.log; tail -15 gpurun_out/r2_pytest_after_fix.log
