#!/bin/bash
# Term recombination (K3): parity tests + a timing of qcut_recombine_launch on resident term values.
set -u
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
timeout 600 python -m pytest tests/test_observable_terms.py -m gpu -x -q 2>&1 | tail -5
python - <<'PY'
import numpy as np, torch, sys
sys.path.insert(0, ".")
from qiskit_addon_cutting_b200 import observable_terms as ot
rng = np.random.default_rng(1)
n_terms, n_rows, per_row = 100_000, 4096, 2048
r = ot.TermRecombination.__new__(ot.TermRecombination)
r.n_rows = n_rows
r.row_start = (np.arange(n_rows + 1) * per_row).astype(np.uint32)
r.index = rng.integers(n_terms, size=n_rows * per_row).astype(np.uint32)
r.coeffs = (rng.normal(size=n_rows * per_row) + 1j * rng.normal(size=n_rows * per_row)).astype(np.complex128)
r._device = None
e = torch.from_numpy(rng.uniform(-1, 1, n_terms)).cuda()
for _ in range(3): out = r.launch(e)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): out = r.launch(e)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
nbytes = r.index.nbytes + r.coeffs.nbytes + 8 * n_terms + 16 * n_rows
print(f"recombine: {n_rows} observables x {per_row} terms: {ms*1e3:.1f} us per launch, {nbytes/ms/1e6:.0f} GB/s of CSR+values")
want = np.array([np.sum(r.coeffs[i*per_row:(i+1)*per_row] * e.cpu().numpy()[r.index[i*per_row:(i+1)*per_row]]) for i in range(4)])
print("first rows (device vs numpy pairwise-sum):", np.abs(out[:4].cpu().numpy() - want).max())
PY
