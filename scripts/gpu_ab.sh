#!/bin/bash
# usage (GPU box): bash scripts/gpu_ab.sh <tag> "<variant names>" "<env settings ;-separated>" [lattices]
tag=${1:-ab}; variants=${2:-"default"}; envs=${3:-""}; lats=${4:-"32,32,32,32 64,64,64,64"}
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
IFS=';' read -ra ENVS <<< "$envs"; [ ${#ENVS[@]} -eq 0 ] && ENVS=("")
for lat in $lats; do for v in $variants; do for e in "${ENVS[@]}"; do
  lib=$PWD/gaugefields.jl_b200/libgfb200.so; [ "$v" != default ] && lib=$PWD/gaugefields.jl_b200/libgfb200_$v.so
  echo "lattice $lat variant $v env [$e]"; env $e GFB200_LIB=$lib timeout 300 $B --lattice $lat 2>>gpurun_out/err_$tag.log | python -c "$S"
done; done; done 2>&1 | tee gpurun_out/ab_$tag.log
tail -5 gpurun_out/err_$tag.log
