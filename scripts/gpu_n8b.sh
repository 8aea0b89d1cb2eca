#!/bin/bash
# usage (8-GPU box): bash scripts/gpu_n8b.sh <tag> -- oracle parity on 4 and 8 ranks, then 64^4 strong scaling at 8 and 4 GPUs (peer-store vs NCCL halos),
# 48^3x96 stout HMC on 8 GPUs, 32^4 flow on 4 GPUs
tag=${1:-n8b}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 4 8; do timeout 300 $T --nproc-per-node $n --master-port $((29700+n)) scripts/dist_check.py 2>&1 | grep -E "dist_check|Error|error|assert" | head -5; done | tee gpurun_out/dist_check_$tag.log
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], (d.get("parity_check") or {}).get("ok"), (d.get("e2e") or {}).get("value"))'
p=29710
run() { p=$((p+1)); echo "N=$1 halo=$2 args [$3]"; GFB200_HALO=$2 timeout 400 $T --nproc-per-node $1 --master-port $p bench.py --gpus $1 --steps 20 --warmup 3 $3 2>>gpurun_out/bench_$tag.err | tee -a gpurun_out/bench_$tag.json | python -c "$S"; }
{
run 8 peer "--no-e2e"
run 8 nccl "--no-e2e --no-parity-check"
run 8 peer ""
run 4 peer "--no-e2e"
run 4 nccl "--no-e2e --no-parity-check"
run 2 peer "--no-e2e"
run 8 peer "--workload stout48 --steps 10"
run 4 peer "--workload flow32 --no-e2e"
run 2 peer "--workload flow32 --no-e2e"
} 2>&1 | tee gpurun_out/ab_$tag.log
tail -5 gpurun_out/bench_$tag.err
