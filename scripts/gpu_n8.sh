#!/bin/bash
# usage (8-GPU box): bash scripts/gpu_n8.sh <tag> -- 64^4 strong scaling at 8 GPUs: which kernel for the six interior slices
tag=${1:-n8}
mkdir -p gpurun_out
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
B="bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e"
echo "k_force_fused everywhere";     GFB200_TMARCH=0 timeout 120 $T --master-port 29571 $B 2>gpurun_out/n8_$tag.err | tee gpurun_out/bench_n8_generic_$tag.json | python -c "$S"
echo "t-marching interior, segments of 3"; GFB200_TMARCH_SEGLEN=3 timeout 120 $T --master-port 29572 $B 2>>gpurun_out/n8_$tag.err | python -c "$S"
echo "t-marching interior (default)"; timeout 120 $T --master-port 29573 $B 2>>gpurun_out/n8_$tag.err | python -c "$S"
