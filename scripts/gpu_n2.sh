#!/bin/bash
# NOTE: NCCL_P2P_USE_CUDA_MEMCPY=1 hangs this bench on the pool's boxes (two 300 s timeouts in round 1): do not add it back.
# usage (2-GPU box): bash scripts/gpu_n2.sh <tag> -- GPU suite, then the 64^4 bench on 2 GPUs: grid policy / SM reservation / NCCL copy-engine sweep
tag=${1:-n2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_full_$tag.log 2>&1; tail -4 gpurun_out/pytest_full_$tag.log
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e"
echo "default (one CTA per item)";          timeout 300 $T --master-port 29511 $B 2>gpurun_out/bench_n2_$tag.err | tee gpurun_out/bench_n2_$tag.json | python -c "$S"
echo "persistent, reserve 16";               GFB200_TMARCH_PERSISTENT=1 GFB200_TMARCH_RESERVE_SMS=16 timeout 300 $T --master-port 29512 $B 2>>gpurun_out/bench_n2_$tag.err | python -c "$S"
echo "default, 64^3x16 per GPU (as at N=4)"; timeout 300 $T --master-port 29515 $B --lattice 64,64,64,32 2>>gpurun_out/bench_n2_$tag.err | python -c "$S"
echo "default, 64^3x8 per GPU (as at N=8)";  timeout 300 $T --master-port 29516 $B --lattice 64,64,64,16 2>>gpurun_out/bench_n2_$tag.err | python -c "$S"
echo "1 GPU 64^4 persistent";     timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline | python -c "$S"
echo "1 GPU 64^4 one CTA per item"; GFB200_TMARCH_PERSISTENT=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline | python -c "$S"
