"""Dynamic instruction mix (executed warp instructions by opcode) and stall samples from an .ncu-rep source page.

    python scripts/ncu_opmix.py gpurun_out/prof_k1.ncu-rep [launch index]
"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
blocks = [(s, (start[j + 1] if j + 1 < len(start) else len(rows))) for j, s in enumerate(start)]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
s, e = blocks[which]
hdr = rows[s + 1]
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0.0, 0.0])
tot_exec = tot_smp = 0.0
for r in rows[s + 2 : e]:
    if len(r) < len(hdr):
        continue
    t = r[ix["Source"]].split()
    if not t:
        continue
    op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
    op = ".".join(op.split(".")[:2]) if op.startswith("IMAD") else op.split(".")[0]
    ex = float(r[ix["Instructions Executed"]] or 0)
    sm = float(r[ix["# Samples"]] or 0)
    agg[op][0] += ex
    agg[op][1] += sm
    tot_exec += ex
    tot_smp += sm
print(f"kernel {rows[s][1]}: {tot_exec/1e6:.1f} M warp instructions, {tot_smp:.0f} samples")
for op, (ex, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:22]:
    print(f"  {op:12s} exec {ex/1e6:9.2f} M ({100*ex/tot_exec:5.1f}%)   samples {100*sm/max(tot_smp,1):5.1f}%")
