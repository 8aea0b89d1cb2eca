#!/bin/bash
# usage: bash scripts/gpu_dram.sh <tag> "<lattices>" -- DRAM bytes + L2 read sectors + duration of the fused MD-step kernel per lattice
tag=${1:-dram}; lats=${2:-"64,64,64,16 64,64,60,16 64,60,64,16 32,32,32,32 32,32,30,32"}
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline"
for lat in $lats; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_tmarch -s 3 -c 1 --csv --log-file gpurun_out/dram_${tag}_$lat.csv $B --lattice $lat > /dev/null 2>&1
  python - gpurun_out/dram_${tag}_$lat.csv $lat <<'PY'
import csv,sys
d={}
for r in csv.reader(open(sys.argv[1])):
    if len(r)>3 and r[-3] not in ("Metric Name",): d[r[-3]]=r[-1]
n=1
for v in sys.argv[2].split(","): n*=int(v)
g=lambda k: float(d.get(k,"nan").replace(",",""))
print(sys.argv[2], "dram read B/site %.0f write %.0f  L2 tex read B/site %.0f hit %.0f  L2 hit rate %.1f%%  time %.3f ms  (%.3f ns/site)"%(g("dram__bytes_read.sum")/n, g("dram__bytes_write.sum")/n, g("lts__t_sectors_srcunit_tex_op_read.sum")*32/n, g("lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum")*32/n, g("lts__t_sector_hit_rate.pct"), g("gpu__time_duration.sum")/1e6, g("gpu__time_duration.sum")/n))
PY
done 2>&1 | tee gpurun_out/dram_$tag.log
