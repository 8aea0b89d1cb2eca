#!/bin/bash
# usage (1-GPU box): bash scripts/gpu_final.sh <tag> -- full GPU suite, the four bench lines, ncu launch list and ncu --set full of the fused kernel at 64^4
tag=${1:-fin}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_full_$tag.log 2>&1; tail -3 gpurun_out/pytest_full_$tag.log
if [ -f gaugefields.jl_b200/libgfb200_e044.so ]; then bash scripts/gpu_ab.sh ${tag}_ab "e044 default e044 default" "" "64,64,64,64"; fi
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_md64_$tag.json 2> gpurun_out/bench_$tag.err; cut -c1-400 gpurun_out/bench_md64_$tag.json
for w in md16 flow32 stout48; do timeout 400 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/bench_${w}_$tag.json 2>> gpurun_out/bench_$tag.err; cut -c1-200 gpurun_out/bench_${w}_$tag.json; done
B="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches64_$tag.csv $B > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_tmarch -s 3 -c 1 -o gpurun_out/prof_ws64_$tag -f $B > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log; tail -3 gpurun_out/bench_$tag.err
