#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
for v in staged:0 direct:4; do
  name=${v%%:*}; flags=${v##*:}
  CMD="python bench.py --workload cfg4_heavy_hex --scale 0.05 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --plan-flags $flags"
  $CMD > $OUT/ncu_plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:qcut_tally_sliced -s 3 -c 1 -f -o $OUT/prof_$name $CMD > $OUT/ncu_$name.log 2>&1
  echo "ncu $name exit $?"
done
