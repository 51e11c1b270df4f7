#!/bin/bash
# Scaling sweep on one box (run with gpurun --gpus 8): N = 1, 2, 4, 8 back to back, kernel-resident + e2e.
set -u
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
nvidia-smi -L | wc -l; free -g | head -2
for n in ${NS:-1 2 4 8}; do
  if [ $n = 1 ]; then
    timeout 900 python bench.py --gpus 1 ${BENCH_ARGS:-} > $OUT/scale_$n.json 2> $OUT/scale_$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n ${BENCH_ARGS:-} > $OUT/scale_$n.json 2> $OUT/scale_$n.err
  fi
  echo "bench N=$n exit $?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/scale_$n.json")); r=d["roofline"]; e=d.get("e2e") or {}
    print(f"  N=$n value {d['value']:.3e} ms/step {d['ms_per_step']:.3f} K1 frac {r['frac']:.3f} k2+allreduce {r['k2_plus_allreduce_ms']:.3f} ms e2e {e.get('value',0):.3e} ({e.get('ms_per_step',0):.0f} ms)")
except Exception as ex:
    print("  no json", ex)
PY
  tail -2 $OUT/scale_$n.err | cut -c1-300
done
