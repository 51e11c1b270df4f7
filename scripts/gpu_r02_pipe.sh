#!/bin/bash
# K1 "pipe" engine (loads of the next step inside the term network) vs the all-at-once engine on cfg4, resident only.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
if [ "${SKIP_PYTEST:-0}" != 1 ]; then timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/r02_pipe_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r02_pipe_pytest.log; fi
run() {
  local name=$1; shift; local envs=$1; shift
  env QCUT_CACHE_DIR=/tmp/qcut_cache_$name $envs timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-per-config --steps 20 --warmup 3 "$@" > $OUT/r02_pipe_$name.json 2> $OUT/r02_pipe_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/r02_pipe_$name.json").read().strip().splitlines()[-1])
    print("$name: frac %.4f  k1 %.4f ms  step %.4f ms  tallies_ok %s clocks %s" % (d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["parity"]["tallies_bit_exact"], d["clocks"]["sm_mhz"]))
except Exception as exc:
    print("$name: no line:", exc)
PY
  tail -1 $OUT/r02_pipe_$name.err | cut -c1-200
}
# arguments: name:ENV=1,ENV=2 ...
for v in "$@"; do
  name=${v%%:*}; envs=${v#*:}
  run "$name" "$(echo "$envs" | tr ',' ' ')"
done
