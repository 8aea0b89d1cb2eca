#!/bin/bash
# usage (1-GPU box): bash scripts/gpu_ncu_ws.sh <tag> -- ncu --set full of the fused MD-step kernel (the one that runs) at 64^4 and 32^4 + launch list at 64^4
tag=${1:-r2ws}
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:k_tmarch -s 3 -c 1 -o gpurun_out/prof_ws64_$tag -f $B > gpurun_out/ncu_ws_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tmarch -s 3 -c 1 -o gpurun_out/prof_ws32_$tag -f $B --lattice 32,32,32,32 >> gpurun_out/ncu_ws_$tag.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches64_$tag.csv $B > /dev/null 2>&1
tail -3 gpurun_out/ncu_ws_$tag.log; ls -la gpurun_out | grep $tag
