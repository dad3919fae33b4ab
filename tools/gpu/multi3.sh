mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 2 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench $N rc $?"; head -c 300 gpurun_out/r2_bench_${N}gpu.json; echo
done
