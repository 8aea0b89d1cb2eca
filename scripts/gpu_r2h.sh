#!/bin/bash
tag=${1:-r2h}
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
run() { echo "lattice $1 env [$2]"; env $2 timeout 300 $B --lattice $1 2>>gpurun_out/err_$tag.log | python -c "$S"; }
{
for e in "GFB200_TMARCH_LAG=0" "GFB200_TMARCH_LAG=1" "GFB200_TMARCH_LAG=2" "GFB200_TMARCH_LAG=3" "GFB200_TMARCH_LAG=6"; do run 64,64,64,64 "$e"; done
for e in "GFB200_TMARCH_LAG=0" "GFB200_TMARCH_LAG=2"; do run 32,32,32,32 "$e"; done
} 2>&1 | tee gpurun_out/ab_$tag.log
for lag in 2; do GFB200_TMARCH_LAG=$lag bash scripts/gpu_dram.sh ${tag}_lag$lag "64,64,64,64"; done
