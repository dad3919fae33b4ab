#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for k in 1 2; do
  echo "== default (pruned)"; timeout 100 python tools/ncu_target.py --spp 1000 --reps 3 | grep rep
  echo "== RT_B200_NO_PRUNE=1"; RT_B200_NO_PRUNE=1 timeout 100 python tools/ncu_target.py --spp 1000 --reps 3 | grep rep
done
