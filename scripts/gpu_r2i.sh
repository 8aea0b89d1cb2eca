#!/bin/bash
# usage (1-GPU box): bash scripts/gpu_r2i.sh <tag> -- non-unitary parity test, then every bench workload once on 1 GPU
tag=${1:-r2i}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_md.py -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -3 gpurun_out/pytest_$tag.log
for wl in flow32 stout48 md16; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_${wl}_$tag.json 2> gpurun_out/bench_${wl}_$tag.err; tail -c 1800 gpurun_out/bench_${wl}_$tag.json; tail -3 gpurun_out/bench_${wl}_$tag.err
done
timeout 900 python bench.py > gpurun_out/bench_md64_$tag.json 2> gpurun_out/bench_md64_$tag.err; tail -c 2500 gpurun_out/bench_md64_$tag.json; tail -3 gpurun_out/bench_md64_$tag.err
