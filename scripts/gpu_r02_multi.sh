#!/bin/bash
# Round 2, N GPUs of one box: NCCL tests of the sharded public API, weak + strong scaling quick legs, staging sweep.
# Usage: gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_r02_multi.sh N'
set -u
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
{ nvidia-smi -L; free -g; nproc; lscpu | grep -E "Model name|Socket|Core|L2|L3|NUMA"; } > $OUT/r02_box_n$N.txt 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
python scripts/host_bw.py --gib 4 > $OUT/r02_host_bw_n$N.json 2>&1; cat $OUT/r02_host_bw_n$N.json
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q > $OUT/r02_pytest_multi_n$N.log 2>&1; echo "pytest exit $?"; tail -15 $OUT/r02_pytest_multi_n$N.log
run() {  # name, extra env, bench args
  local name=$1; shift; local envs=$1; shift
  env $envs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 \
      bench.py --gpus $N --no-cpu-baseline --no-per-config "$@" > $OUT/r02_n${N}_$name.json 2> $OUT/r02_n${N}_$name.err
  echo "$name exit $?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/r02_n${N}_$name.json").read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print("  value %.3g  ms %.3f  frac %.3f  scaling %s | e2e %.4g (%.1f ms, %.1f GB/s per rank)" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["scaling"], e.get("value", 0), e.get("ms_per_step", 0), e.get("host_gbs_per_rank", 0)))
except Exception as exc:
    print("  no line:", exc)
PY
  tail -2 $OUT/r02_n${N}_$name.err
}
# staging sweep (end to end, a quarter of the tuples per rank; same throughput regime, a quarter of the time)
run sweep_auto "X=1" --scale 0.25 --steps 5 --e2e-steps 3
run sweep_ring64_t8 "QCUT_STAGING_MB=64 QCUT_STAGING_THREADS=8" --scale 0.25 --steps 5 --e2e-steps 3
run sweep_ring8_t2 "QCUT_STAGING_MB=8 QCUT_STAGING_THREADS=2" --scale 0.25 --steps 5 --e2e-steps 3
run sweep_ring4_t2 "QCUT_STAGING_MB=4 QCUT_STAGING_THREADS=2" --scale 0.25 --steps 5 --e2e-steps 3
run sweep_ring16_t4 "QCUT_STAGING_MB=16 QCUT_STAGING_THREADS=4" --scale 0.25 --steps 5 --e2e-steps 3
run sweep_ring2_t1 "QCUT_STAGING_MB=2 QCUT_STAGING_THREADS=1" --scale 0.25 --steps 5 --e2e-steps 3
# scaling legs: weak (97 000 tuples per GPU) and strong (97 000 tuples in total), full size, default policy
run weak "X=1" --steps 20 --e2e-steps 3
run strong "X=1" --scaling strong --steps 20 --e2e-steps 3
run cfg5_weak "X=1" --workload cfg5_molecular --steps 20 --e2e-steps 3
run cfg5_strong "X=1" --workload cfg5_molecular --scaling strong --steps 20 --e2e-steps 3
