mkdir -p gpurun_out
timeout 600 python tools/ab_global.py --spp 200 > gpurun_out/r2_ab17_global.log 2>&1; cat gpurun_out/r2_ab17_global.log
RT_B200_LIB=rust-tracing_b200/csrc/librt_b200_base.so timeout 600 python tools/ab_global.py --spp 200 > gpurun_out/r2_ab17_global_base.log 2>&1; cat gpurun_out/r2_ab17_global_base.log
timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -q -k "specialised or layout" 2>&1 | tail -3
timeout 600 python tools/fuzz_render.py --rich --seeds 40 2>&1 | tail -2
