#!/bin/bash
# Sweep launch shapes / staging of the specialised K1 kernel on a scaled cfg4 (kernel-resident timing only).
set -u
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or width or many_terms" > $OUT/pytest_tune.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_tune.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload ${WL:-cfg4_heavy_hex} --scale ${SCALE:-0.3} --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --plan-flags ${FLAGS:-0} > $OUT/tune_$name.json 2> $OUT/tune_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/tune_$name.json")); r=d["roofline"]
    print(f"$name: K1 {r['kernel_ms']:.3f} ms  {r['achieved']:.0f} GB/s  frac {r['frac']:.3f}  step {d['ms_per_step']:.3f} ms")
except Exception as e:
    print("$name: failed", e); print(open("$OUT/tune_$name.err").read()[-600:])
PY
}
FLAGS=0 run base QCUT_JIT_PREFETCH=0
FLAGS=0 run pf1 QCUT_JIT_PREFETCH=1
FLAGS=0 run pf2 QCUT_JIT_PREFETCH=2
FLAGS=4 run staged QCUT_JIT_PREFETCH=0
FLAGS=4 run staged_128x4 QCUT_JIT_THREADS=128 QCUT_JIT_MINBLOCKS=4
