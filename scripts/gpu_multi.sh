#!/bin/bash
# Multi-GPU check (run with gpurun --gpus N): NCCL test + bench at 1..N ranks.
set -u
N=${N:-2}
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q > $OUT/pytest_multi.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_multi.log
for n in 1 $N; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 ${BENCH_ARGS:-} > $OUT/scale_$n.json 2> $OUT/scale_$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 10 --warmup 3 ${BENCH_ARGS:-} > $OUT/scale_$n.json 2> $OUT/scale_$n.err
  fi
  echo "bench N=$n exit $?"; tail -c 1500 $OUT/scale_$n.json; echo; tail -3 $OUT/scale_$n.err
done
