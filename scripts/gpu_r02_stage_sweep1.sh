#!/bin/bash
# Staging sweep on ONE rank (e2e only, a quarter of cfg4): ring size x threads x store type, host copy ceiling beside it.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
python scripts/host_bw.py --gib 4 | tee $OUT/r02_host_bw_n1.json
run() {
  local name=$1; shift; local envs=$1; shift
  env $envs timeout 600 python bench.py --warm-cache --no-cpu-baseline --no-per-config --scale 0.25 --steps 3 --e2e-steps 5 > $OUT/r02_sw1_$name.json 2> $OUT/r02_sw1_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/r02_sw1_$name.json").read().strip().splitlines()[-1]); e = d["e2e"]
    print("$name: e2e %.1f ms  %.1f GB/s  (copy ceiling %.1f GB/s, frac %.2f)  first call %.1f ms" % (e["ms_per_step"], e["host_gbs_per_rank"], e["host_copy_ceiling"]["copy_gbs"], e["host_copy_frac"], d["cold_start"]["first_call_ms"]))
except Exception as exc:
    print("$name: no line:", exc)
PY
  tail -1 $OUT/r02_sw1_$name.err | cut -c1-200
}
run default "X=1"
run ring256_nt0 "QCUT_STAGING_MB=256 QCUT_STAGING_NT=0"
run ring32_t8 "QCUT_STAGING_MB=32 QCUT_STAGING_THREADS=8"
run ring16_t8 "QCUT_STAGING_MB=16 QCUT_STAGING_THREADS=8"
run ring8_t8 "QCUT_STAGING_MB=8 QCUT_STAGING_THREADS=8 QCUT_STAGING_MIN_SLOT_KB=256"
run ring16_t4 "QCUT_STAGING_MB=16 QCUT_STAGING_THREADS=4"
run ring64_t8_nt1 "QCUT_STAGING_MB=64 QCUT_STAGING_THREADS=8 QCUT_STAGING_NT=1"
run default_again "X=1"
python scripts/host_bw.py --gib 4
