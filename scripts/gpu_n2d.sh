#!/bin/bash
# usage (2-GPU box): bash scripts/gpu_n2d.sh <tag> -- t-marching tests, then 8-slice slabs on 2 GPUs: staggered start on/off, peer stores suppressed (timing only)
tag=${1:-n2d}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tmarch.py tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_$tag.log
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], (d.get("parity_check") or {}).get("ok"))'
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 40 --warmup 3 --no-e2e --no-cpu-baseline"
p=29650
for rep in 1 2; do for lat in 64,64,64,16 64,64,64,32; do for e in "GFB200_TMARCH_STAGGER=1" "GFB200_TMARCH_STAGGER=0" "GFB200_PEER_NOSTORE=1"; do
  p=$((p+1)); extra=""; [ "$e" = "GFB200_PEER_NOSTORE=1" ] && extra="--no-parity-check"
  echo "lattice $lat $e"; env $e timeout 200 $T --master-port $p $B $extra --lattice $lat 2>>gpurun_out/err_$tag.log | python -c "$S"
done; done; done 2>&1 | tee gpurun_out/ab_$tag.log
tail -3 gpurun_out/err_$tag.log
