set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "negative_radius" > gpurun_out/r2_pytest_new2.log 2>&1; echo "rc $?" >> gpurun_out/r2_pytest_new2.log; tail -40 gpurun_out/r2_pytest_new2.log
