set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -q -x -k "specialised" > gpurun_out/r2_pytest_specialised.log 2>&1; echo "rc $?" >> gpurun_out/r2_pytest_specialised.log; tail -5 gpurun_out/r2_pytest_specialised.log
timeout 600 python tools/fuzz_render.py --seeds 32 > gpurun_out/r2_fuzz_render.log 2>&1; echo "rc $?" >> gpurun_out/r2_fuzz_render.log; tail -8 gpurun_out/r2_fuzz_render.log
timeout 600 python tools/fuzz_parity.py --seeds 16 --rays 32768 > gpurun_out/r2_fuzz_parity.log 2>&1; echo "rc $?" >> gpurun_out/r2_fuzz_parity.log; tail -3 gpurun_out/r2_fuzz_parity.log
