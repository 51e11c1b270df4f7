#!/bin/bash
# Round 2, one gpurun call: build check, GPU parity tests, smoke, the default bench line (headline + per_config).
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_r02_check.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
{ nvidia-smi; free -g; nproc; lscpu | head -30; } > $OUT/box.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r02_pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r02_pytest_gpu.log
tail -15 $OUT/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/smoke.log
timeout 1200 python bench.py --steps 20 --warmup 3 > $OUT/r02_bench_default.json 2> $OUT/r02_bench_default.err; echo "bench default exit $?"
cut -c1-3000 $OUT/r02_bench_default.json; tail -5 $OUT/r02_bench_default.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02_bench_reference.json 2> $OUT/r02_bench_reference.err; cut -c1-900 $OUT/r02_bench_reference.json
