#!/bin/bash
# usage (1-GPU box): bash scripts/gpu_ncu_other.sh <tag> -- wall times and per-kernel ncu metrics of the paths outside the three benchmarks, 32^4
tag=${1:-other}
mkdir -p gpurun_out
timeout 300 python scripts/other_kernels.py 32,32,32,32 --time 2>&1 | tee gpurun_out/other_times_$tag.log
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed_pipe_fp64.sum"
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/ncu_other32_$tag.csv python scripts/other_kernels.py 32,32,32,32 > /dev/null 2>&1
python scripts/ncu_summary.py gpurun_out/ncu_other32_$tag.csv 1048576 | tee gpurun_out/ncu_other32_$tag.md | cut -c1-200
