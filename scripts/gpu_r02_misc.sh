#!/bin/bash
# Round 2, one GPU: run-time-mask kernel on the many-group shapes, small-call profile, cold start, ncu captures per kernel.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
quick() {
  local name=$1; shift; local envs=$1; shift
  env $envs timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-per-config --steps 20 --warmup 3 "$@" > $OUT/r02_misc_$name.json 2> $OUT/r02_misc_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/r02_misc_$name.json").read().strip().splitlines()[-1])
    print("$name: frac %.4f  k1 %.4f ms  step %.4f ms  kernels %s tallies_ok %s" % (d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["roofline"]["k1_kernels_per_launch"], d["parity"]["tallies_bit_exact"]))
except Exception as exc:
    print("$name: no line:", exc)
PY
  tail -1 $OUT/r02_misc_$name.err | cut -c1-200
}
# 1. run-time-mask kernel (ahead-of-time compiled, no NVRTC) vs the specialised kernels
for wl in cfg5_molecular cfg3_wide_local cfg2_wire cfg1_tutorial cfg4_heavy_hex; do
  quick rt_$wl "X=1" --workload $wl --plan-flags 1
done
quick jit_cfg5 "X=1" --workload cfg5_molecular
# 2. cold start: plan creation + first call with an empty kernel cache
for wl in cfg4_heavy_hex cfg5_molecular cfg1_tutorial; do
  rm -rf /tmp/qcut_cold_cache
  QCUT_CACHE_DIR=/tmp/qcut_cold_cache timeout 300 python - <<PY 2>&1 | tail -2
import sys, time
sys.path[:0] = ["."]
import torch, workloads as W
from qiskit_addon_cutting_b200 import engine
w = W.get_workload("$wl", 0.01 if "$wl" != "cfg1_tutorial" else 1.0)
tables = W.build_tables(w)
data = W.device_data(tables, 1)
torch.cuda.synchronize()
t = time.perf_counter(); plan = engine.Plan(tables.groups, tables.masks); cold = time.perf_counter() - t
t = time.perf_counter(); plan2 = engine.Plan(tables.groups, tables.masks); warm = time.perf_counter() - t
t = time.perf_counter(); plan3 = engine.Plan(tables.groups, tables.masks, flags=1); rt = time.perf_counter() - t
print("cold start $wl: plan with empty cache %.3f s, same plan from the disk cache %.3f s, run-time-mask plan (no NVRTC) %.4f s, kernels %d" % (cold, warm, rt, len(set(plan.kernels()))))
PY
done
# 3. where a small call spends its host time
timeout 300 python scripts/profile_small.py cfg1_tutorial 300 > $OUT/r02_profile_small_cfg1.txt 2>&1; head -45 $OUT/r02_profile_small_cfg1.txt
timeout 300 python scripts/profile_small.py cfg2_wire 100 > $OUT/r02_profile_small_cfg2.txt 2>&1; head -30 $OUT/r02_profile_small_cfg2.txt
# 4. ncu --set full, one launch per kernel (each only after the plain command exited 0)
cap() {  # name, kernel regex, skip, bench args...
  local name=$1; shift; local rx=$1; shift; local skip=$1; shift
  local CMD="python bench.py --no-e2e --no-cpu-baseline --no-per-config --steps 2 --warmup 3 $*"
  $CMD > $OUT/r02_ncu_${name}_plain.json 2> $OUT/r02_ncu_${name}_plain.err && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $OUT/r02_prof_$name $CMD > $OUT/r02_ncu_$name.log 2>&1
  echo "ncu $name exit $?"
}
cap k1_cfg4 qcut_tally_sliced 6 --workload cfg4_heavy_hex --scale 0.25
# cfg5: all 13 kernels of one qcut_tally_launch (8 single-group + 5 multi-group), no source import (size)
CMD5="python bench.py --no-e2e --no-cpu-baseline --no-per-config --steps 2 --warmup 3 --workload cfg5_molecular --scale 0.5"
$CMD5 > $OUT/r02_ncu_k1_cfg5_plain.json 2> $OUT/r02_ncu_k1_cfg5_plain.err && \
timeout 900 ncu --set full --clock-control none -k regex:qcut_tally_sliced -s 39 -c 13 -f -o $OUT/r02_prof_k1_cfg5 $CMD5 > $OUT/r02_ncu_k1_cfg5.log 2>&1
echo "ncu k1_cfg5 exit $?"
cap k1_cfg3_dense qcut_tally_sliced 6 --workload cfg3_wide_dense
cap k1_rt tally_sliced_rt_kernel 3 --workload cfg4_heavy_hex --scale 0.1 --plan-flags 1
cap k1_wide qcut_tally_sliced 3 --workload cfg4u_heavy_hex_unpartitioned --scale 0.1
cap k2_stream reduce_stream_kernel 3 --workload cfg4_heavy_hex --scale 0.25
cap literal literal_expvals_kernel 0 --workload cfg4_heavy_hex --scale 0.02
ls -la $OUT/*.ncu-rep | tail -10
