#!/bin/bash
# Quick iteration: parity tests + kernel-resident benches (+ optional ncu capture of the K1 kernel).
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
if [ "${TESTS:-1}" = "1" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
fi
for wl in ${WORKLOADS:-cfg4_heavy_hex cfg3_wide_dense cfg5_molecular cfg3_wide_local cfg2_wire}; do
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
  echo "bench $wl exit $?"; python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$wl.json"))
    r=d["roofline"]
    print(f"  {d['value']:.3e} shot-terms/s  {d['ms_per_step']:.3f} ms/step  K1 {r['kernel_ms']:.3f} ms  {r['achieved']:.0f} GB/s  frac {r['frac']:.3f}  launches/step {d['gpu_launches_per_step']}")
except Exception as e:
    print("  (no json)", e)
PY
  tail -2 $OUT/bench_$wl.err
done
if [ "${NCU:-1}" = "1" ]; then
CMD="python bench.py --workload ${NCU_WL:-cfg4_heavy_hex} --scale ${NCU_SCALE:-0.05} --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > $OUT/ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:${NCU_K:-qcut_tally_sliced} -s 3 -c 1 -f -o $OUT/prof_k1 $CMD > $OUT/ncu_full.log 2>&1
echo "ncu exit $?"
fi
