#!/bin/bash
# Round 2 scaling legs on N GPUs of one box (the driver's own SCALE run covers weak cfg4 with e2e at every N).
# Usage: gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_r02_scale.sh N'
set -u
N=${1:-8}
OUT=gpurun_out
mkdir -p $OUT
{ nvidia-smi -L; free -g; nproc; lscpu | grep -E "Model name|Socket|Core|L2|L3|NUMA"; } > $OUT/r02_box_n$N.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
python scripts/host_bw.py --gib 4 > $OUT/r02_host_bw_n$N.json 2>&1; cat $OUT/r02_host_bw_n$N.json
run() {  # name, bench args
  local name=$1; shift
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --no-cpu-baseline --no-per-config "$@" > $OUT/r02_scale_n${N}_$name.json 2> $OUT/r02_scale_n${N}_$name.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29643 \
      bench.py --gpus $N --no-cpu-baseline --no-per-config "$@" > $OUT/r02_scale_n${N}_$name.json 2> $OUT/r02_scale_n${N}_$name.err
  fi
  echo "$name exit $?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/r02_scale_n${N}_$name.json").read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("  value %.4g  ms %.4f  frac %.3f  k1 %.4f ms  k2+ar %.4f ms  scaling %s | e2e %.4g (%.1f ms, %.1f GB/s per rank, copy frac %s)" % (
        d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["k2_plus_allreduce_ms"], d["scaling"],
        e.get("value", 0), e.get("ms_per_step", 0), e.get("host_gbs_per_rank", 0), e.get("host_copy_frac")))
except Exception as exc:
    print("  no line:", exc)
PY
  tail -2 $OUT/r02_scale_n${N}_$name.err | cut -c1-300
}
if [ "$N" = "1" ]; then
  run strong_cfg4 --scaling strong --steps 20 --no-e2e
  run strong_cfg4_s65536 --scaling strong --shots 65536 --steps 5 --no-e2e
  run strong_cfg5 --workload cfg5_molecular --scaling strong --steps 20 --no-e2e
else
  run strong_cfg4 --scaling strong --steps 20 --e2e-steps 3
  run strong_cfg4_s65536 --scaling strong --shots 65536 --steps 10 --no-e2e
  run strong_cfg5 --workload cfg5_molecular --scaling strong --steps 20 --no-e2e
  run weak_cfg5 --workload cfg5_molecular --steps 20 --no-e2e
fi
