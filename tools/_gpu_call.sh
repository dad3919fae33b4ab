#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_hits.py tests/test_gpu_render.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/ab.py '[{"RT_B200_NO_PRUNE":"1"},{},{"RT_B200_SLAB_EXIT":"6"},{"RT_B200_MIN_BLOCKS":"5"}]' 8,6,0,7,3 128 > gpurun_out/ab20.log 2>&1
cat gpurun_out/ab20.log
