#!/bin/bash
# Round 2, second pass after the switch kernels: cfg5 strong + weak and cfg4 strong at N GPUs (resident; the end-to-end
# legs at N > 1 are host bound and were measured before), NCCL tests of the sharded API at N = 2.
# Usage: gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_r02_scale2.sh N'
set -u
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q > $OUT/r02_pytest_multi_n$N.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/r02_pytest_multi_n$N.log
fi
run() {  # name, bench args
  local name=$1; shift
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --no-cpu-baseline --no-per-config --no-e2e "$@" > $OUT/r02_scale_n${N}_$name.json 2> $OUT/r02_scale_n${N}_$name.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29643 \
      bench.py --gpus $N --no-cpu-baseline --no-per-config --no-e2e "$@" > $OUT/r02_scale_n${N}_$name.json 2> $OUT/r02_scale_n${N}_$name.err
  fi
  echo "$name exit $?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/r02_scale_n${N}_$name.json").read().strip().splitlines()[-1])
    print("  value %.4g  ms %.4f  frac %.3f  k1 %.4f ms  k2+ar %.4f ms  scaling %s" % (
        d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["k2_plus_allreduce_ms"], d["scaling"]))
except Exception as exc:
    print("  no line:", exc)
PY
  tail -1 $OUT/r02_scale_n${N}_$name.err | cut -c1-300
}
run strong_cfg5 --workload cfg5_molecular --scaling strong --steps 20
if [ "$N" != "1" ]; then run weak_cfg5 --workload cfg5_molecular --steps 20; fi
run strong_cfg4 --scaling strong --steps 20
