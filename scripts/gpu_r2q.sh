#!/bin/bash
# usage: bash scripts/gpu_r2q.sh <tag> -- stout / baseline-lattice tests, then the stout48 and flow32 bench lines on one GPU
tag=${1:-r2q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_flow_stout.py tests/test_gpu_baseline_lattices.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_$tag.log
for w in stout48 flow32; do
  timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline 2>>gpurun_out/err_$tag.log | tee gpurun_out/bench_${w}_$tag.json | cut -c1-600
done
