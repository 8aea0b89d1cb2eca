#!/bin/bash
# usage (2-GPU box): bash scripts/gpu_n2c.sh <tag> -- full GPU suite incl. multi-GPU parity, then the default bench at N = 2 with the e2e leg
tag=${1:-n2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_full_$tag.log 2>&1; tail -6 gpurun_out/pytest_full_$tag.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $T --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 3 2>>gpurun_out/bench_$tag.err | tee gpurun_out/bench_$tag.json | python -c 'import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["parity_check"]["ok"], d["e2e"]["value"], d["e2e"]["phases_rank0"], d["e2e"]["pcie_gbs_rank0"], d["e2e"]["host_binding_rank0"])'
tail -3 gpurun_out/bench_$tag.err
