#!/bin/bash
# usage (1-GPU box): bash scripts/gpu_r2f.sh <tag> -- tile-block shape (L2) and warp-stagger A/B of the t-marching kernel + DRAM bytes at 64^4
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tmarch.py tests/test_gpu_baseline_lattices.py -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -2 gpurun_out/pytest_$tag.log
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline"
S='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
run() { echo "lattice $1 env [$2]"; env $2 timeout 300 $B --lattice $1 2>>gpurun_out/err_$tag.log | python -c "$S"; }
{
for e in "X=0" "GFB200_TMARCH_BY=16 GFB200_TMARCH_BZ=1" "GFB200_TMARCH_BY=2 GFB200_TMARCH_BZ=9" "GFB200_TMARCH_BY=4 GFB200_TMARCH_BZ=4" "GFB200_TMARCH_BY=4 GFB200_TMARCH_BZ=5" "GFB200_TMARCH_BY=6 GFB200_TMARCH_BZ=3" "GFB200_TMARCH_WS=0" "GFB200_TMARCH_STAGGER=2000" "GFB200_TMARCH_STAGGER=5000"; do run 64,64,64,64 "$e"; done
for e in "X=0" "GFB200_TMARCH_BY=8 GFB200_TMARCH_BZ=1" "GFB200_TMARCH_WS=0" "GFB200_TMARCH_STAGGER=1000" "GFB200_TMARCH_STAGGER=2500" "GFB200_TMARCH_STAGGER=5000"; do run 32,32,32,32 "$e"; done
} 2>&1 | tee gpurun_out/ab_$tag.log
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:k_tmarch -s 3 -c 1 --csv --log-file gpurun_out/dram64_$tag.csv $B > /dev/null 2>&1
tail -3 gpurun_out/dram64_$tag.csv
tail -5 gpurun_out/err_$tag.log
