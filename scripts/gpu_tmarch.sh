#!/bin/bash
# usage (GPU box): bash scripts/gpu_tmarch.sh <tag> [variants]  -- parity of the t-marching kernel, A/B bench against variants, ncu
tag=${1:-tm}
variants=${2:-""}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tmarch.py tests/test_gpu_parity_md.py -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
tail -5 gpurun_out/pytest_$tag.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
for lat in 32,32,32,32 64,64,64,64; do
  echo "lattice $lat default";  timeout 300 $B --lattice $lat 2>gpurun_out/err_$tag.log | python -c "$S"
  for v in $variants; do
    echo "lattice $lat variant $v"; GFB200_LIB=$PWD/gaugefields.jl_b200/libgfb200_$v.so timeout 300 $B --lattice $lat 2>>gpurun_out/err_$tag.log | python -c "$S"
  done
done 2>&1 | tee gpurun_out/ab_$tag.log
[ -n "$SKIP_NCU" ] || timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --lattice 32,32,32,32 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1
[ -n "$SKIP_NCU" ] || timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tmarch -s 3 -c 1 -o gpurun_out/prof_tmarch_$tag -f \
    python bench.py --lattice 32,32,32,32 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
