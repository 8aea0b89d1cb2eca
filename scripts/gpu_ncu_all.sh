#!/bin/bash
# usage (1-GPU box): bash scripts/gpu_ncu_all.sh <tag> -- per-kernel ncu metrics (duration, DRAM bytes, FP64 pipe, LSU, registers) of every kernel of the
# flow32, stout (32^4) and md (32^4) workloads; summarised by scripts/ncu_summary.py into profiles/
tag=${1:-r2all}
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_read.sum,smsp__inst_executed_pipe_fp64.sum"
ncu --metrics $M --clock-control none -c 40 --csv --log-file gpurun_out/ncu_flow32_$tag.csv python bench.py --workload flow32 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --metrics $M --clock-control none -c 60 --csv --log-file gpurun_out/ncu_stout32_$tag.csv python bench.py --workload stout48 --lattice 32,32,32,32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --metrics $M --clock-control none -c 30 --csv --log-file gpurun_out/ncu_md32_$tag.csv python bench.py --workload md64 --lattice 32,32,32,32 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*$tag*
