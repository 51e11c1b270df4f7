#!/bin/bash
# One gpurun call: build check, GPU parity tests, smoke, benches, ncu launch list + one full capture.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
{ nvidia-smi; free -g; nproc; } > $OUT/box.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/smoke.log
for wl in cfg4_heavy_hex cfg3_wide_dense cfg5_molecular cfg3_wide_local cfg2_wire cfg1_tutorial; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-e2e --cpu-seconds 4 > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err
  echo "bench $wl exit $?"; cat $OUT/bench_$wl.json | cut -c1-700; tail -2 $OUT/bench_$wl.err
done
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench default exit $?"
cat $OUT/bench_default.json; tail -3 $OUT/bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>&1; cat $OUT/bench_reference.json | cut -c1-600
# ncu: launch list, then one full capture of the dominant kernel (only after the plain command exited 0)
CMD="python bench.py --workload cfg4_heavy_hex --scale 0.05 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > $OUT/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:qcut_tally_sliced -s 3 -c 2 -o $OUT/prof_k1 $CMD > $OUT/ncu_full.log 2>&1
echo "ncu exit $?"; tail -5 $OUT/launches.csv
