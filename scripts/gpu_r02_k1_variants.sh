#!/bin/bash
# K1 variants on cfg4 (resident only): split engine x CTA shape.  One gpurun call, ~25 s per variant.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or every_row_width or many_terms" > $OUT/r02_k1v_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r02_k1v_pytest.log
run() {
  local name=$1; shift; local envs=$1; shift
  env QCUT_CACHE_DIR=/tmp/qcut_cache_$name $envs timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-per-config --steps 20 --warmup 3 "$@" > $OUT/r02_k1v_$name.json 2> $OUT/r02_k1v_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/r02_k1v_$name.json").read().strip().splitlines()[-1])
    print("$name: frac %.4f  k1 %.4f ms  step %.4f ms  tallies_ok %s clocks %s" % (d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["parity"]["tallies_bit_exact"], d["clocks"]["sm_mhz"]))
except Exception as exc:
    print("$name: no line:", exc)
PY
  tail -1 $OUT/r02_k1v_$name.err | cut -c1-200
}
run base_256x2 "QCUT_JIT_SPLIT=0"
run split_256x2 "QCUT_JIT_SPLIT=1 QCUT_JIT_THREADS=256 QCUT_JIT_MINBLOCKS=2"
run split_256x3 "QCUT_JIT_SPLIT=1 QCUT_JIT_THREADS=256 QCUT_JIT_MINBLOCKS=3"
run split_192x3 "QCUT_JIT_SPLIT=1 QCUT_JIT_THREADS=192 QCUT_JIT_MINBLOCKS=3"
run split_128x5 "QCUT_JIT_SPLIT=1 QCUT_JIT_THREADS=128 QCUT_JIT_MINBLOCKS=5"
run split_128x4 "QCUT_JIT_SPLIT=1 QCUT_JIT_THREADS=128 QCUT_JIT_MINBLOCKS=4"
run split_128x6 "QCUT_JIT_SPLIT=1 QCUT_JIT_THREADS=128 QCUT_JIT_MINBLOCKS=6"
run split2_256x3 "QCUT_JIT_SPLIT=1 QCUT_JIT_SPLIT_PHASES=2 QCUT_JIT_THREADS=256 QCUT_JIT_MINBLOCKS=3"
run base_192x3 "QCUT_JIT_SPLIT=0 QCUT_JIT_THREADS=192 QCUT_JIT_MINBLOCKS=3"
