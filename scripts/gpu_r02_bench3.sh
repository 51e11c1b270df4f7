#!/bin/bash
# Slab-engine GPU tests, the un-partitioned 127-qubit extra workload with both wide-row engines, then the driver's bench command.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "slab or every_row_width or rows_of_17" > $OUT/r02_pytest_slab.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r02_pytest_slab.log
bash scripts/gpu_r02_wl_variants.sh cfg4u_heavy_hex_unpartitioned slab:0: wide:0:QCUT_JIT_SLAB=0 | tail -4
START=$(date +%s)
timeout 1500 python bench.py > $OUT/r02_bench_default.json 2> $OUT/r02_bench_default.err; echo "bench default exit $? in $(( $(date +%s) - START )) s"
tail -3 $OUT/r02_bench_default.err | cut -c1-300
python scripts/bench_tables.py $OUT/r02_bench_default.json | cut -c1-420
