#!/bin/bash
# End-to-end rate vs size of the pinned staging ring, at N ranks (run with gpurun --gpus N).
set -u
N=${N:-2}
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
nproc; free -g | head -2; lscpu | grep -i "model name\|socket\|L3\|numa" | head -8
for mb in ${SIZES:-256 64 32}; do
  for n in ${NS:-1 $N}; do
    if [ $n = 1 ]; then
      QCUT_STAGING_MB=$mb timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/sweep_${mb}_$n.json 2> $OUT/sweep_${mb}_$n.err
    else
      QCUT_STAGING_MB=$mb timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > $OUT/sweep_${mb}_$n.json 2> $OUT/sweep_${mb}_$n.err
    fi
    python - <<PY
import json
try:
    d=json.load(open("$OUT/sweep_${mb}_$n.json")); e=d["e2e"]
    print(f"ring $mb MiB N=$n: e2e {e['value']:.3e} shot-terms/s, {e['ms_per_step']:.0f} ms per call per rank")
except Exception as ex:
    print("ring $mb N=$n: no json", ex)
PY
  done
done
