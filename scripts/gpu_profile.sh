#!/bin/bash
# Round-end evidence: GPU tests, default bench, ncu launch list of the same command, one full capture of K1.
set -u
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > $OUT/profile_plain.json 2> $OUT/profile_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:qcut_tally_sliced -s 6 -c 2 -f -o $OUT/prof_k1 $CMD > $OUT/ncu_full.log 2>&1
echo "ncu exit $?"
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench exit $?"; cat $OUT/bench_default.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>&1; cut -c1-400 $OUT/bench_reference.json
for wl in cfg3_wide_dense cfg3_wide_local cfg5_molecular cfg2_wire cfg1_tutorial cfg4u_heavy_hex_unpartitioned; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --cpu-seconds 3 > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "bench $wl exit $?"
done
bash scripts/gpu_terms.sh > $OUT/terms.log 2>&1; tail -3 $OUT/terms.log
bash scripts/gpu_pcie.sh > $OUT/pcie.log 2>&1; cat $OUT/pcie.log
