#!/bin/bash
# usage (2-GPU box): bash scripts/gpu_n2e.sh <tag> -- where the per-step overhead of a slab run goes: peer stores / flag wait suppressed (timing only)
tag=${1:-n2e}
mkdir -p gpurun_out
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 40 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check"
p=29750
echo "single GPU 64,64,64,16"; python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu-baseline --lattice 64,64,64,16 2>>gpurun_out/err_$tag.log | python -c "$S"
for lat in 64,64,64,32; do for e in "A=1" "GFB200_PEER_NOSTORE=1" "GFB200_HALO_NOWAIT=1" "GFB200_PEER_NOSTORE=1 GFB200_HALO_NOWAIT=1"; do
  p=$((p+1)); echo "lattice $lat $e"; env $e timeout 200 $T --master-port $p $B --lattice $lat 2>>gpurun_out/err_$tag.log | python -c "$S"
done; done 2>&1 | tee gpurun_out/ab_$tag.log
echo "single GPU 64,64,64,16"; python bench.py --steps 40 --warmup 3 --no-e2e --no-cpu-baseline --lattice 64,64,64,16 2>>gpurun_out/err_$tag.log | python -c "$S"
tail -3 gpurun_out/err_$tag.log
