set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "layout_switches or rich or generated or specialised" > gpurun_out/r2_pytest_new.log 2>&1; echo "rc $?" >> gpurun_out/r2_pytest_new.log; tail -30 gpurun_out/r2_pytest_new.log
