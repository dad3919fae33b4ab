mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_scatter.py -x -q > gpurun_out/r2_pytest_multi.log 2>&1; tail -8 gpurun_out/r2_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; python -c "
import json;d=json.load(open('gpurun_out/r2_bench_2gpu.json'));print(d['value'],d['e2e'],d['config'],d['check'])"; tail -3 gpurun_out/r2_bench_2gpu.err
