#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
LF=$GRAFT_REPO_ROOT/rust-tracing_b200/csrc/librt_b200_fold.so
RT_B200_LIB=$LF timeout 300 python -m pytest tests/test_gpu_render.py tests/test_gpu_wavefront.py -x -q -m gpu 2>&1 | tail -3
V='[{},{"RT_B200_LIB":"'$LF'"},{},{"RT_B200_LIB":"'$LF'"}]'
timeout 400 python tools/ab.py "$V" 8,6,0,7 256 > gpurun_out/ab23.log 2>&1
sed "s#$GRAFT_REPO_ROOT/rust-tracing_b200/csrc/##" gpurun_out/ab23.log
