#!/bin/bash
# usage (8-GPU box): bash scripts/gpu_n8c.sh <tag> -- dist_check on 8 ranks, then the default bench at N = 8 (with e2e) and N = 4
tag=${1:-n8c}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $T --nproc-per-node 8 --master-port 29801 scripts/dist_check.py 2>&1 | grep -E "dist_check|Error|error|assert" | head -5 | tee gpurun_out/dist_check_$tag.log
nvidia-smi topo -m > gpurun_out/topo_$tag.txt 2>&1; lscpu | grep -i -E "numa|socket|model name" >> gpurun_out/topo_$tag.txt
for n in 8 4; do
  timeout 240 $T --nproc-per-node $n --master-port $((29810+n)) bench.py --gpus $n --steps 20 --warmup 3 2>>gpurun_out/bench_$tag.err | tee -a gpurun_out/bench_$tag.json | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["parity_check"]["ok"], d["e2e"]["value"], d["e2e"]["phases_rank0"], d["e2e"]["pcie_gbs_rank0"], d["e2e"]["host_binding_rank0"])'
done 2>&1 | tee gpurun_out/ab_$tag.log
tail -3 gpurun_out/bench_$tag.err; cat gpurun_out/topo_$tag.txt | head -30
