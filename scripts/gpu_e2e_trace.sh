#!/bin/bash
# Per-call timing of the public API on cfg4 with the staging trace (where do slow calls lose time?).
set -u
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
for cfg in ${SIZES:-256 64}; do
IFS=: read mb th nt <<< "$cfg"; th=${th:-8}; nt=${nt:-0}
echo "== ring $mb MiB threads $th streaming-stores $nt"
QCUT_STAGING_MB=$mb QCUT_STAGING_THREADS=$th QCUT_STAGING_NT=$nt python - <<'PY'
import os, time, numpy as np, torch, sys
sys.path.insert(0, ".")
import workloads as W
from qiskit_addon_cutting_b200 import engine, reconstruct_expectation_values
w = W.get_workload("cfg4_heavy_hex", 1.0)
tab = W.build_tables(w)
data = W.device_data(tab, 1)
host = np.empty(tab.data_bytes, dtype=np.uint8); torch.from_numpy(host).copy_(data[: tab.data_bytes])
res = W.host_results(w, tab, host)
coeffs = w.coefficients()[:w.n_coeffs]
pairs = W.coefficient_pairs(coeffs)
keep_resident = os.environ.get("KEEP_RESIDENT", "1") == "1"
if not keep_resident:
    del data
print(f"ring {os.environ['QCUT_STAGING_MB']} MiB, {os.environ['QCUT_STAGING_THREADS']} gather threads x 2 feeders, resident copy kept: {keep_resident}")
import gc
gclog = []
def _cb(phase, info):
    if phase == "start": gclog.append([info["generation"], time.perf_counter(), None])
    else: gclog[-1][2] = time.perf_counter()
gc.callbacks.append(_cb)
marks = {}
def _timed(cls, name):
    orig = getattr(cls, name)
    def wrapper(*a, **k):
        t = time.perf_counter()
        try: return orig(*a, **k)
        finally: marks[cls.__name__ + "." + name] = marks.get(cls.__name__ + "." + name, 0.0) + time.perf_counter() - t
    setattr(cls, name, wrapper)
_timed(engine.DeviceBatch, "run"); _timed(engine.Schedule, "__del__"); _timed(engine.Schedule, "__init__")
import qiskit_addon_cutting_b200.cutting_reconstruction as cr
_orig_stage = engine.stage_v2
def _stage(*a, **k):
    t = time.perf_counter()
    try: return _orig_stage(*a, **k)
    finally: marks["stage_v2"] = time.perf_counter() - t; marks["stage_v2_start"] = t
engine.stage_v2 = _stage
for rep in range(int(os.environ.get("CALLS", "10"))):
    gclog.clear(); marks.clear()
    t0 = time.perf_counter(); got = reconstruct_expectation_values(res, pairs, w.observables); t1 = time.perf_counter()
    print("   pre-stage %.0f ms, stage_v2 %.0f, after %.0f | %s | gc %s" % (
        1e3 * (marks["stage_v2_start"] - t0), 1e3 * marks["stage_v2"], 1e3 * (t1 - marks["stage_v2_start"] - marks["stage_v2"]),
        {k: round(1e3 * v, 1) for k, v in marks.items() if "." in k}, [(g, round(1e3 * (e - s))) for g, s, e in gclog]))
    tr = engine.last_stage_trace
    blocks = " ".join(f"{b/1e9:.2f}GB:{s*1e3:.0f}-{e*1e3:.0f}" for b, s, e in tr["blocks"])
    print(f"call {rep}: {1e3*(t1-t0):.0f} ms  walk {1e3*tr['walk_s']:.0f}  tables {1e3*tr['tables_s']:.0f}  wait {1e3*tr['wait_s']:.0f} | {blocks}")
PY
done
