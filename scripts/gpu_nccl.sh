#!/bin/bash
# usage (2-GPU box): bash scripts/gpu_nccl.sh <tag> -- multi-GPU parity, then small-slab step time with and without the two-row halo exchange
tag=${1:-nccl}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_multi_$tag.log 2>&1; tail -3 gpurun_out/pytest_multi_$tag.log
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
p=29540
for lat in 64,64,64,16 64,64,64,32 64,64,64,64; do
  for su3 in 1 0; do
    p=$((p+1))
    echo "lattice $lat two-row halo $su3"
    GFB200_HALO_SU3=$su3 timeout 150 $T --master-port $p bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --lattice $lat 2>>gpurun_out/nccl_$tag.err | python -c "$S"
  done
done
