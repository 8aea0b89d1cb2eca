#!/bin/bash
# usage (1-GPU box): bash scripts/gpu_r2g.sh <tag> -- pacing of neighbouring copy streams (GFB200_TMARCH_LAG): parity, timing A/B, DRAM bytes at 64^4
tag=${1:-r2g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tmarch.py tests/test_gpu_baseline_lattices.py -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -2 gpurun_out/pytest_$tag.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
run() { echo "lattice $1 env [$2]"; env $2 timeout 300 $B --lattice $1 2>>gpurun_out/err_$tag.log | python -c "$S"; }
{
for e in "GFB200_TMARCH_LAG=0" "GFB200_TMARCH_LAG=1" "GFB200_TMARCH_LAG=2" "GFB200_TMARCH_LAG=4" "GFB200_TMARCH_LAG=8"; do run 64,64,64,64 "$e"; done
for e in "GFB200_TMARCH_LAG=0" "GFB200_TMARCH_LAG=1" "GFB200_TMARCH_LAG=2" "GFB200_TMARCH_LAG=4"; do run 32,32,32,32 "$e"; done
} 2>&1 | tee gpurun_out/ab_$tag.log
for lag in 0 1 2 4; do GFB200_TMARCH_LAG=$lag bash scripts/gpu_dram.sh ${tag}_lag$lag "64,64,64,64"; done
tail -5 gpurun_out/err_$tag.log
