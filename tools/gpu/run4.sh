mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bvh_build.py tests/test_gpu_jpeg.py tests/test_gpu_render.py -q -s -k "bvh or jpeg or noise_textures" > gpurun_out/r2_pytest4.log 2>&1; tail -30 gpurun_out/r2_pytest4.log
