#!/bin/bash
# usage (on the GPU box, via gpurun): bash scripts/gpu_profile.sh <tag>
# 1) gpu parity tests  2) bench line  3) ncu launch list  4) ncu --set full of the fused force kernel (k_tmarch_fused at the bench lattice)
tag=${1:-r1}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
tail -3 gpurun_out/pytest_$tag.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -1 gpurun_out/bench_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tmarch_fused -s 3 -c 2 -o gpurun_out/prof_force_$tag -f \
    python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_update_links -s 1 -c 1 -o gpurun_out/prof_links_$tag -f \
    python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline >> gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out
