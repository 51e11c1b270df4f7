#!/bin/bash
# ncu --set full of the cfg4 K1 kernel for engine variants: args name:ENV=..,ENV=..
set -u
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
for v in "$@"; do
  name=${v%%:*}; envs=$(echo "${v#*:}" | tr ',' ' ')
  CMD="python bench.py --workload cfg4_heavy_hex --scale 0.25 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config"
  env QCUT_CACHE_DIR=/tmp/qc_$name $envs $CMD > $OUT/ncu_plain_$name.log 2>&1 && \
  env QCUT_CACHE_DIR=/tmp/qc_$name $envs ncu --set full --clock-control none --import-source on -k regex:qcut_tally_sliced -s 6 -c 1 -f -o $OUT/r02_prof_v_$name $CMD > $OUT/ncu_v_$name.log 2>&1
  echo "ncu $name exit $?"
done
