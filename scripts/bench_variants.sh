# usage: bash scripts/bench_variants.sh "<suffix list>" [lattice]   (tuning variants built by build.py with GFB200_VARIANT)
LAT=${2:-32,32,32,32}
for v in ${1:-""}; do
  [ "$v" = "default" ] && v=""
  echo "variant [$v]"
  GFB200_LIB=$PWD/gaugefields.jl_b200/libgfb200$v.so python bench.py --lattice $LAT --steps 20 --warmup 3 --no-e2e --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
