#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
if [ "${TESTS:-1}" = "1" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log
fi
[ "${BENCH:-1}" = "1" ] && timeout 900 python bench.py --steps 10 --warmup 3 --cpu-seconds 4 --e2e-steps 3 > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_default.json"))
print("value %.3e  ms/step %.3f  frac %.3f  e2e %.3e (%.1f ms/step)  cpu %.3e" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["cpu_baseline"]["value"]))
PY
tail -3 $OUT/bench_default.err
python - <<'PY'
# where does the end-to-end time go?
import time, numpy as np, torch, sys
sys.path.insert(0, ".")
import workloads as W
from qiskit_addon_cutting_b200 import engine, _lib, reconstruct_expectation_values
w = W.get_workload("cfg4_heavy_hex", 1.0)
tab = W.build_tables(w)
data = W.device_data(tab, 1)
host = np.empty(tab.data_bytes, dtype=np.uint8); torch.from_numpy(host).copy_(data[: tab.data_bytes])
res = W.host_results(w, tab, host)
colls = w.collections(); coeffs = w.coefficients()[:w.n_coeffs]
lst = [res[l] for l in w.observables]
del data
for rep in range(3):
    t0 = time.perf_counter(); tables, arrays = engine.build_host_tables(lst, colls, coeffs, 0, w.n_coeffs, w.n_obs); t1 = time.perf_counter()
    dev = torch.empty(tables.data_bytes + 128, dtype=torch.uint8, device="cuda")
    pinned = engine._pinned_buffer(engine._STAGING_BYTES)
    torch.cuda.synchronize(); ta = time.perf_counter()
    engine.check(_lib.load().qcut_stage_h2d(_lib.np_ptr(arrays.src), _lib.np_ptr(arrays.nbytes), _lib.np_ptr(arrays.dst), len(arrays),
                 pinned.data_ptr(), pinned.numel(), dev.data_ptr(), 0, engine._stream_ptr()), "stage")
    torch.cuda.synchronize(); tb = time.perf_counter()
    del dev
    t2 = time.perf_counter()
    batch = engine.stage_v2(lst, colls, coeffs, 0, w.n_coeffs, w.n_obs); torch.cuda.synchronize(); t3 = time.perf_counter()
    out = batch.run(); torch.cuda.synchronize(); t4 = time.perf_counter()
    del batch
    t5 = time.perf_counter(); got = reconstruct_expectation_values(res, W.coefficient_pairs(coeffs), w.observables); t6 = time.perf_counter()
    print(f"rep {rep}: walk+tables {t1-t0:.3f}s  stage_h2d alone {tb-ta:.3f}s ({tables.data_bytes/(tb-ta)/1e9:.1f} GB/s)  stage_v2 {t3-t2:.3f}s  kernels {t4-t3:.4f}s  public API {t6-t5:.3f}s")
import cProfile, pstats, io
pairs = W.coefficient_pairs(coeffs)
for rep in range(2):
    pr = cProfile.Profile(); t5 = time.perf_counter(); pr.enable()
    got = reconstruct_expectation_values(res, pairs, w.observables)
    pr.disable(); t6 = time.perf_counter()
    print(f"profiled public API call {t6-t5:.3f}s")
sio = io.StringIO(); pstats.Stats(pr, stream=sio).sort_stats("tottime").print_stats(18); print(sio.getvalue()[:3500])
PY
