#!/bin/bash
# usage (8-GPU box): bash scripts/gpu_n8.sh <tag> -- 64^4 strong scaling at 8 GPUs: halo-exchange variants
tag=${1:-n8}
mkdir -p gpurun_out
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
B="bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e"
echo "two-row halo";                 timeout 120 $T --master-port 29561 $B 2>gpurun_out/n8_$tag.err | tee gpurun_out/bench_n8_$tag.json | python -c "$S"
echo "full halo";                    GFB200_HALO_SU3=0 timeout 120 $T --master-port 29562 $B 2>>gpurun_out/n8_$tag.err | python -c "$S"
echo "two-row halo, 32 p2p channels"; NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32 timeout 120 $T --master-port 29563 $B 2>>gpurun_out/n8_$tag.err | python -c "$S"
echo "two-row halo, 16 SMs reserved"; GFB200_TMARCH_RESERVE_SMS=16 timeout 120 $T --master-port 29564 $B 2>>gpurun_out/n8_$tag.err | python -c "$S"
