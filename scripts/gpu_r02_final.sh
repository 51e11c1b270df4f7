#!/bin/bash
# Round 2 evidence run (one GPU): GPU tests, smoke, the driver's bench commands (ours + reference arm), the ncu launch
# list of the bench command, and ncu --set full summaries of the kernels whose reports are too big to bring back.
set -u
OUT=gpurun_out
mkdir -p $OUT
{ nvidia-smi; free -g; nproc; lscpu | grep -E "Model name|Socket|Core|L2|L3|NUMA"; } > $OUT/r02_box_final.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/r02_pytest_gpu_final.log 2>&1; echo "pytest exit $?"; grep -E "first call|S=|at scale|passed|failed|Error" $OUT/r02_pytest_gpu_final.log | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r02_bench_default.json 2> $OUT/r02_bench_default.err; echo "bench default exit $?"

python scripts/bench_tables.py $OUT/r02_bench_default.json | cut -c1-400
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/r02_bench_reference.json 2> $OUT/r02_bench_reference.err; cut -c1-700 $OUT/r02_bench_reference.json
timeout 300 python scripts/profile_small.py cfg1_tutorial 300 > $OUT/r02_profile_small_cfg1.txt 2>&1; head -1 $OUT/r02_profile_small_cfg1.txt
timeout 300 python scripts/profile_small.py cfg2_wire 100 > $OUT/r02_profile_small_cfg2.txt 2>&1; head -1 $OUT/r02_profile_small_cfg2.txt
CMD="python bench.py --warm-cache --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config"
$CMD > $OUT/r02_launches_plain.json 2> $OUT/r02_launches_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"qcut|reduce_|literal|tally" -c 200 --csv --log-file $OUT/r02_cfg4_bench_launches.csv $CMD > $OUT/r02_ncu_launches.log 2>&1
echo "ncu launch list exit $?"
cap() {  # name, kernel regex, skip, count, bench args...
  local name=$1; shift; local rx=$1; shift; local skip=$1; shift; local cnt=$1; shift
  local CMD="python bench.py --warm-cache --no-e2e --no-cpu-baseline --no-per-config --steps 2 --warmup 3 $*"
  $CMD > $OUT/r02_ncu_${name}_plain.json 2> $OUT/r02_ncu_${name}_plain.err && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -f -o /tmp/r02_prof_$name $CMD > $OUT/r02_ncu_$name.log 2>&1
  echo "ncu $name exit $?"
  python scripts/ncu_summary.py /tmp/r02_prof_$name.ncu-rep > $OUT/r02_${name}_ncu_summary.txt 2>&1
  python scripts/ncu_opmix.py /tmp/r02_prof_$name.ncu-rep > $OUT/r02_${name}_opcode_mix.txt 2>&1
}
cap k1_cfg4 qcut_tally_sliced 6 2 --workload cfg4_heavy_hex
cap k1_cfg5 qcut_tally_sliced 12 4 --workload cfg5_molecular --scale 0.5
cap k1_cfg3_local qcut_tally_sliced 6 2 --workload cfg3_wide_local
cap k2_stream reduce_stream_kernel 3 1 --workload cfg4_heavy_hex
python scripts/ncu_traffic.py /tmp/r02_prof_k1_cfg4.ncu-rep $OUT/r02_k1_traffic.json cfg4_heavy_hex 1.0
du -sh $OUT
