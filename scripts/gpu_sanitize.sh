#!/bin/bash
# compute-sanitizer memcheck on the small parity cases (one tool per gpurun call).
set -u
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or width or zero or subsystems" > $OUT/sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file $OUT/memcheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or width or zero or subsystems" > $OUT/sanitize_run.log 2>&1
echo "memcheck exit $?"; tail -3 $OUT/sanitize_run.log; grep -E "ERROR SUMMARY|Invalid|out of bounds" $OUT/memcheck.log | head -10
