#!/bin/bash
# usage: bash scripts/gpu_rs.sh <tag> -- tmarch parity tests, A/B of the round barrier (GFB200_TMARCH_ROUNDSYNC=0/1) at 64^4 and 32^4, DRAM bytes at 64^4
tag=${1:-rs}
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_tmarch.py tests/test_gpu_parity_md.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_$tag.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
for rep in 1 2; do for lat in 64,64,64,64 32,32,32,32; do for rs in 0 1; do
  echo "lattice $lat ROUNDSYNC=$rs"; GFB200_TMARCH_ROUNDSYNC=$rs timeout 120 $B --lattice $lat 2>>gpurun_out/err_$tag.log | python -c "$S"
done; done; done 2>&1 | tee gpurun_out/ab_$tag.log
for rs in 0 1; do GFB200_TMARCH_ROUNDSYNC=$rs bash scripts/gpu_dram.sh ${tag}_rs$rs "64,64,64,64"; done
tail -3 gpurun_out/err_$tag.log
