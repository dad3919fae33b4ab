set -x

mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 300 python tools/ncu_target.py --spp 32 --reps 3 > gpurun_out/r2_first.log 2>&1; echo "rc=$?" >> gpurun_out/r2_first.log
tail -5 gpurun_out/r2_first.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest1.log
tail -30 gpurun_out/r2_pytest1.log
