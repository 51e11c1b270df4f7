#!/bin/bash
# Round 2 check #3: GPU tests, small-call timings, default bench with cold-start blocks, ncu summaries made ON the box.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 1200 python -m pytest tests -m gpu -x -q -s > $OUT/r02_pytest_gpu3.log 2>&1; echo "pytest exit $?"; grep -E "first call|S=|passed|failed|Error" $OUT/r02_pytest_gpu3.log | tail -12
timeout 300 python scripts/profile_small.py cfg1_tutorial 300 > $OUT/r02_profile_small_cfg1.txt 2>&1; head -3 $OUT/r02_profile_small_cfg1.txt | cut -c1-200
timeout 300 python scripts/profile_small.py cfg2_wire 100 > $OUT/r02_profile_small_cfg2.txt 2>&1; head -3 $OUT/r02_profile_small_cfg2.txt | cut -c1-200
timeout 1500 python bench.py --steps 20 --warmup 3 > $OUT/r02_bench_default.json 2> $OUT/r02_bench_default.err; echo "bench default exit $?"
tail -3 $OUT/r02_bench_default.err | cut -c1-300
python scripts/bench_tables.py $OUT/r02_bench_default.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_default.json").read().strip().splitlines()[-1])
for r in [d] + d.get("per_config", []):
    print(r["config"]["workload"], json.dumps(r.get("cold_start")), "e2e path", (r.get("e2e") or {}).get("staging_path"))
print(json.dumps(d.get("dedup")))
PY
# ncu --set full, one or a few launches per kernel; summaries are produced here and the big reports are not kept
cap() {  # name, kernel regex, skip, count, bench args...
  local name=$1; shift; local rx=$1; shift; local skip=$1; shift; local cnt=$1; shift
  local CMD="python bench.py --warm-cache --no-e2e --no-cpu-baseline --no-per-config --steps 2 --warmup 3 $*"
  $CMD > $OUT/r02_ncu_${name}_plain.json 2> $OUT/r02_ncu_${name}_plain.err && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o /tmp/r02_prof_$name $CMD > $OUT/r02_ncu_$name.log 2>&1
  echo "ncu $name exit $?"
  python scripts/ncu_summary.py /tmp/r02_prof_$name.ncu-rep > $OUT/r02_${name}_ncu_summary.txt 2>&1
  python scripts/ncu_opmix.py /tmp/r02_prof_$name.ncu-rep > $OUT/r02_${name}_opcode_mix.txt 2>&1
  if [ $(stat -c %s /tmp/r02_prof_$name.ncu-rep 2>/dev/null || echo 99999999) -lt 8000000 ]; then cp /tmp/r02_prof_$name.ncu-rep $OUT/; fi
}
cap k1_cfg4 qcut_tally_sliced 6 2 --workload cfg4_heavy_hex
cap k1_cfg5 qcut_tally_sliced 39 13 --workload cfg5_molecular --scale 0.5
cap k1_cfg3_dense qcut_tally_sliced 6 2 --workload cfg3_wide_dense
cap k1_cfg3_local qcut_tally_sliced 27 9 --workload cfg3_wide_local
cap k1_rt tally_sliced_rt_kernel 3 1 --workload cfg4_heavy_hex --scale 0.1 --plan-flags 1
cap k1_generic tally_generic_kernel 3 1 --workload cfg4_heavy_hex --scale 0.02 --plan-flags 2
cap k1_wide qcut_tally_sliced 3 1 --workload cfg4u_heavy_hex_unpartitioned --scale 0.1
cap k2_stream reduce_stream_kernel 3 1 --workload cfg4_heavy_hex
cap literal literal_expvals_kernel 0 1 --workload cfg4_heavy_hex --scale 0.05
python scripts/ncu_traffic.py /tmp/r02_prof_k1_cfg4.ncu-rep $OUT/r02_k1_traffic.json cfg4_heavy_hex 1.0
du -sh $OUT
