#!/bin/bash
# Resident-only K1 numbers for a workload under plan flags / env variants.  args: workload name:flags:ENV=..,ENV=.. ...
set -u
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
wl=$1; shift
for v in "$@"; do
  name=$(echo $v | cut -d: -f1); flags=$(echo $v | cut -d: -f2); envs=$(echo $v | cut -d: -f3- | tr ',' ' ')
  env QCUT_CACHE_DIR=/tmp/qc_${wl}_$name $envs timeout 600 python bench.py --workload $wl --no-e2e --no-cpu-baseline --no-per-config --steps 20 --warmup 3 --plan-flags $flags > $OUT/r02_wl_${wl}_$name.json 2> $OUT/r02_wl_${wl}_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/r02_wl_${wl}_$name.json").read().strip().splitlines()[-1])
    print("$wl $name: frac %.4f  k1 %.4f ms  step %.4f ms  tallies_ok %s clocks %s" % (d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["parity"]["tallies_bit_exact"], d["clocks"]["sm_mhz"]))
except Exception as exc:
    print("$wl $name: no line:", exc)
PY
  tail -1 $OUT/r02_wl_${wl}_$name.err | cut -c1-200
done
