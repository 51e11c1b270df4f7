"""Compile the specialised kernel of a workload's group here (NVRTC, no GPU) and summarise its SASS.

    python scripts/sass_stats.py cfg4_heavy_hex [group]
"""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import workloads as W  # noqa: E402
from qiskit_addon_cutting_b200 import _lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4_heavy_hex"
group = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tab = W.build_tables(W.get_workload(name, 0.001))
src = _lib.generate_source(tab.groups, tab.masks, group)
Path("/tmp/qcut_k.cu").write_text(src)
Path("/tmp/qcut_k.cubin").write_bytes(_lib.compile_source(src))
print(subprocess.run(["cuobjdump", "-res-usage", "/tmp/qcut_k.cubin"], capture_output=True, text=True).stdout.strip().splitlines()[-1])
sass = subprocess.run(["cuobjdump", "-sass", "/tmp/qcut_k.cubin"], capture_output=True, text=True).stdout
lines = []
for ln in sass.splitlines():
    ln = ln.strip()
    if re.match(r"/\*[0-9a-f]{4,}\*/", ln):
        body = ln.split("*/", 1)[1].split("/*")[0].strip().rstrip(";").strip()
        if body:
            lines.append(body)
Path("/tmp/qcut_k.sass").write_text("\n".join(lines))
ops = Counter()
for b in lines:
    tok = b.split()
    op = tok[1] if tok[0].startswith("@") else tok[0]
    ops[op.split(".")[0]] += 1
print(len(lines), "instructions;", ", ".join(f"{k} {v}" for k, v in ops.most_common(16)))
popc = [i for i, b in enumerate(lines) if "POPC" in b]
ldg = [i for i, b in enumerate(lines) if "LDG.E.128" in b]
print("LDG.128 at", ldg[:3], "...", ldg[-3:], "| POPC from", popc[0], "to", popc[-1])
