#!/bin/bash
# K2 stream-kernel parameters (libraries built beforehand into build_variants/): step - K1 time on cfg4.
set -u
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
for so in build_variants/libqcut_*.so; do
  name=$(basename $so .so)
  QCUT_B200_LIB=$PWD/$so timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-per-config --steps 20 --warmup 3 > $OUT/r02_k2v_$name.json 2> $OUT/r02_k2v_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/r02_k2v_$name.json").read().strip().splitlines()[-1])
    print("$name: step %.4f ms  k1 %.4f  step-k1 %.4f ms  checksum %r tallies %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["k2_plus_allreduce_ms"], d["checksum"], d["parity"]["tallies_bit_exact"]))
except Exception as exc:
    print("$name: no line:", exc)
PY
done
