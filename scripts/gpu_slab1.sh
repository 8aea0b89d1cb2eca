#!/bin/bash
# usage: bash scripts/gpu_slab1.sh <tag> -- slab-shaped lattices on ONE GPU (what a rank of a 4/8-GPU run computes): round barrier off / forced
tag=${1:-slab1}
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
for lat in 64,64,64,8 64,64,64,16 64,64,64,32; do for rs in 0 2; do
  echo "lattice $lat ROUNDSYNC=$rs"; GFB200_TMARCH_ROUNDSYNC=$rs timeout 120 $B --lattice $lat 2>>gpurun_out/err_$tag.log | python -c "$S"
done; done 2>&1 | tee gpurun_out/ab_$tag.log
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full_$tag.log 2>&1; tail -4 gpurun_out/pytest_full_$tag.log
