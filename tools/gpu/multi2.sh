mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2_pytest_multi.log 2>&1; echo "rc $?" >> gpurun_out/r2_pytest_multi.log; tail -4 gpurun_out/r2_pytest_multi.log
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench rc $?"; head -c 900 gpurun_out/r2_bench_${N}gpu.json
python -c "import __graft_entry__ as g; g.smoke()"
