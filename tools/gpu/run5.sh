mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2_pytest5.log 2>&1; echo "rc $?" >> gpurun_out/r2_pytest5.log; tail -15 gpurun_out/r2_pytest5.log
