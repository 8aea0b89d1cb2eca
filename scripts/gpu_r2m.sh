#!/bin/bash
tag=${1:-r2m}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_flow_stout.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
for v in default cl2 cl4; do
  lib=$PWD/gaugefields.jl_b200/libgfb200.so; [ "$v" != default ] && lib=$PWD/gaugefields.jl_b200/libgfb200_$v.so
  echo "variant $v"; GFB200_LIB=$lib timeout 300 python bench.py --workload flow32 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["energy_density_ms"])'
done 2>&1 | tee gpurun_out/ab_$tag.log
