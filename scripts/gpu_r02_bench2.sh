#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
START=$(date +%s)
timeout 1500 python bench.py > $OUT/r02_bench_default.json 2> $OUT/r02_bench_default.err; echo "bench default exit $? in $(( $(date +%s) - START )) s"
tail -3 $OUT/r02_bench_default.err | cut -c1-300
python scripts/bench_tables.py $OUT/r02_bench_default.json | head -9 | cut -c1-300
bash scripts/gpu_r02_launch_list.sh
