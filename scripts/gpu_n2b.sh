#!/bin/bash
# usage (2-GPU box): bash scripts/gpu_n2b.sh <tag> -- GPU suite incl. multi-GPU parity, then peer-store vs NCCL halo exchange on 2 GPUs
tag=${1:-n2b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full_$tag.log 2>&1; tail -6 gpurun_out/pytest_full_$tag.log
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"].get("kernel"))'
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
p=29600
for lat in 64,64,64,64 64,64,64,32 64,64,64,16; do for halo in peer nccl; do
  p=$((p+1)); echo "lattice $lat halo $halo"; GFB200_HALO=$halo timeout 300 $T --master-port $p $B --lattice $lat 2>>gpurun_out/bench_n2_$tag.err | tee -a gpurun_out/bench_n2_$tag.json | python -c "$S"
done; done 2>&1 | tee gpurun_out/ab_$tag.log
tail -5 gpurun_out/bench_n2_$tag.err
