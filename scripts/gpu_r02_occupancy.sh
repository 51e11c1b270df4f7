#!/bin/bash
# CTAs per SM of the narrow-row / few-term K1 kernels (cfg5, cfg3 local): resident only.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
run() {
  local name=$1; shift; local envs=$1; shift
  env $envs timeout 400 python bench.py --warm-cache --no-e2e --no-cpu-baseline --no-per-config --steps 20 --warmup 3 "$@" > $OUT/r02_occ_$name.json 2> $OUT/r02_occ_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/r02_occ_$name.json").read().strip().splitlines()[-1])
    print("$name: frac %.4f  k1 %.4f ms  step %.4f ms  kernels %s tallies_ok %s" % (d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["roofline"]["k1_kernels_per_launch"], d["parity"]["tallies_bit_exact"]))
except Exception as exc:
    print("$name: no line:", exc)
PY
  tail -1 $OUT/r02_occ_$name.err | cut -c1-200
}
for wl in cfg5_molecular cfg3_wide_local; do
  run ${wl}_mb2 "QCUT_JIT_MULTI_MINBLOCKS=2 QCUT_JIT_AUTO_MINBLOCKS=0" --workload $wl
  run ${wl}_mb3 "QCUT_JIT_MULTI_MINBLOCKS=3 QCUT_JIT_AUTO_MINBLOCKS=0" --workload $wl
  run ${wl}_mb4 "QCUT_JIT_MULTI_MINBLOCKS=4 QCUT_JIT_AUTO_MINBLOCKS=0" --workload $wl
  run ${wl}_auto "X=1" --workload $wl
done
run cfg2_auto "X=1" --workload cfg2_wire
run cfg1_auto "X=1" --workload cfg1_tutorial
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "golden or every_row_width or three_kernels" > $OUT/r02_occ_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/r02_occ_pytest.log
