#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -30 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q > gpurun_out/r02_pytest_multi_n4.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r02_pytest_multi_n4.log
bash scripts/gpu_r02_scale.sh 4
