#!/bin/bash
# usage: bash scripts/gpu_ab2.sh <tag> "<variants>" -- timing A/B (64^4 and 32^4, 2 repeats) + DRAM bytes at 64^4 per library variant
tag=${1:-ab2}; variants=${2:-"default"}
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
for rep in 1 2; do for lat in 64,64,64,64 32,32,32,32; do for v in $variants; do
  lib=$PWD/gaugefields.jl_b200/libgfb200.so; [ "$v" != default ] && lib=$PWD/gaugefields.jl_b200/libgfb200_$v.so
  echo "lattice $lat variant $v"; GFB200_LIB=$lib timeout 300 $B --lattice $lat 2>>gpurun_out/err_$tag.log | python -c "$S"
done; done; done 2>&1 | tee gpurun_out/ab_$tag.log
for v in $variants; do
  lib=$PWD/gaugefields.jl_b200/libgfb200.so; [ "$v" != default ] && lib=$PWD/gaugefields.jl_b200/libgfb200_$v.so
  echo "variant $v"; GFB200_LIB=$lib bash scripts/gpu_dram.sh ${tag}_$v "64,64,64,64"
done
