#!/bin/bash
# Per-kernel durations (ncu launch list, cold-cache, serialised) of one workload under env variants: wl name:ENV=..,..
set -u
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
wl=$1; shift
for v in "$@"; do
  name=${v%%:*}; envs=$(echo "${v#*:}" | tr ',' ' ')
  CMD="python bench.py --workload $wl --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config"
  env QCUT_CACHE_DIR=/tmp/qc_ll_$name $envs $CMD > $OUT/ll_plain_$name.log 2>&1
  env QCUT_CACHE_DIR=/tmp/qc_ll_$name $envs ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,launch__grid_size,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:qcut_tally -c 200 --csv --log-file $OUT/r02_ll_${wl}_$name.csv $CMD > $OUT/ll_ncu_$name.log 2>&1
  echo "ncu $name exit $?"
done
