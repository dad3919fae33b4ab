set -x
mkdir -p gpurun_out
timeout 900 python tools/fuzz_render.py --rich --seeds 96 > gpurun_out/r2_fuzz_render_rich.log 2>&1; echo "rc $?" >> gpurun_out/r2_fuzz_render_rich.log; grep -v ": ok " gpurun_out/r2_fuzz_render_rich.log | tail -30
timeout 900 python tools/fuzz_parity.py --rich --seeds 64 --rays 32768 > gpurun_out/r2_fuzz_parity_rich.log 2>&1; echo "rc $?" >> gpurun_out/r2_fuzz_parity_rich.log; grep -v ": ok " gpurun_out/r2_fuzz_parity_rich.log | cut -c1-400 | tail -30
timeout 900 python tools/fuzz_render.py --seeds 64 > gpurun_out/r2_fuzz_render.log 2>&1; echo "rc $?" >> gpurun_out/r2_fuzz_render.log; grep -v ": ok " gpurun_out/r2_fuzz_render.log | tail
