#!/bin/bash
# usage: bash scripts/gpu_gen.sh <tag> "<variants>" -- general-action parity tests, then wall times of the secondary paths per library variant
tag=${1:-gen}; variants=${2:-"default"}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_general_action.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_$tag.log
for v in $variants; do
  lib=$PWD/gaugefields.jl_b200/libgfb200.so; [ "$v" != default ] && lib=$PWD/gaugefields.jl_b200/libgfb200_$v.so
  echo "variant $v"; GFB200_LIB=$lib timeout 200 python scripts/other_kernels.py 32,32,32,32 --time 2>&1 | head -4
done 2>&1 | tee gpurun_out/ab_$tag.log
