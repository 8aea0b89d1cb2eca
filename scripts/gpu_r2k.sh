#!/bin/bash
tag=${1:-r2k}
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
run() { echo "lattice $1 env [$2]"; env $2 timeout 300 $B --lattice $1 2>>gpurun_out/err_$tag.log | python -c "$S"; }
{
for e in "X=0" "GFB200_TMARCH_SEGLEN=32" "GFB200_TMARCH_SEGLEN=22" "GFB200_TMARCH_SEGLEN=16" "GFB200_TMARCH_SEGLEN=8" "X=1"; do run 64,64,64,64 "$e"; done
} 2>&1 | tee gpurun_out/ab_$tag.log
for sl in 16 32; do GFB200_TMARCH_SEGLEN=$sl bash scripts/gpu_dram.sh ${tag}_seg$sl "64,64,64,64"; done
