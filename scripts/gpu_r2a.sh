#!/bin/bash
# usage (GPU box): bash scripts/gpu_r2a.sh <tag>  -- parity + A/B of the t-marching variants (env switches, one build)
tag=${1:-r2a}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tmarch.py tests/test_gpu_parity_md.py tests/test_gpu_parity_flow_stout.py -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
tail -5 gpurun_out/pytest_$tag.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
for lat in 32,32,32,32 64,64,64,64; do
  for ws in 0 1; do for sw in 0 1; do
    echo "lattice $lat ws=$ws swizzle=$sw"; GFB200_TMARCH_WS=$ws GFB200_TMARCH_SWIZZLE=$sw timeout 300 $B --lattice $lat 2>>gpurun_out/err_$tag.log | python -c "$S"
  done; done
done 2>&1 | tee gpurun_out/ab_$tag.log
tail -20 gpurun_out/err_$tag.log
