#!/bin/bash
# usage: bash scripts/gpu_stoutab.sh <tag> "<variants>" -- stout48 step parts per library variant (occupancy of the back-prop kernels)
tag=${1:-stab}; variants=${2:-"default"}
mkdir -p gpurun_out
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["roofline"]["parts_ms"])'
for v in $variants; do
  lib=$PWD/gaugefields.jl_b200/libgfb200.so; [ "$v" != default ] && lib=$PWD/gaugefields.jl_b200/libgfb200_$v.so
  echo "variant $v"; GFB200_LIB=$lib timeout 200 python bench.py --workload stout48 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline 2>>gpurun_out/err_$tag.log | python -c "$S"
done 2>&1 | tee gpurun_out/ab_$tag.log
tail -3 gpurun_out/err_$tag.log
