#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed or layout_cache or mid_size or background or small_public" > $OUT/r02_pytest_new.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r02_pytest_new.log
START=$(date +%s)
timeout 1500 python bench.py > $OUT/r02_bench_default.json 2> $OUT/r02_bench_default.err; echo "bench default exit $? in $(( $(date +%s) - START )) s"
tail -3 $OUT/r02_bench_default.err | cut -c1-300
python scripts/bench_tables.py $OUT/r02_bench_default.json | cut -c1-420
