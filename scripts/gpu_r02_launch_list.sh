#!/bin/bash
# ncu launch list (gpu__time_duration.sum, no clock control) of OUR kernels in the bench command; the torch kernels that
# generate the synthetic data are outside the timed region and are filtered out by name.
set -u
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -30 $OUT/build.log; exit 1; }
CMD="python bench.py --warm-cache --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config"
$CMD > $OUT/r02_launches_plain.json 2> $OUT/r02_launches_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"qcut|reduce_|literal|tally" -c 200 --csv --log-file $OUT/r02_cfg4_bench_launches.csv $CMD > $OUT/r02_ncu_launches.log 2>&1
echo "ncu launch list exit $?"
python - <<'PY'
import csv
from collections import Counter, defaultdict
rows = [r for r in csv.reader(open("gpurun_out/r02_cfg4_bench_launches.csv")) if len(r) > 10 and r[0].isdigit()]
c, t = Counter(), defaultdict(float)
for r in rows:
    name = r[4].split("(")[0][:48]
    c[name] += 1
    t[name] += float(r[-1])
total = sum(t.values())
with open("gpurun_out/r02_cfg4_bench_launches_summary.txt", "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -k regex:qcut|reduce_|literal|tally  python bench.py --warm-cache --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-per-config\n")
    f.write("(3 warm-up + 2 timed steps, then one reference-rounding pass (literal_expvals_kernel + sequential reduce) twice and one parity tally; per-launch times are cold-cache and serialised)\n")
    for k, v in c.most_common():
        line = f"{v:4d} launches  {t[k] / 1e6:10.3f} ms  {100 * t[k] / total:5.1f} %  {k}"
        f.write(line + "\n")
        print(line)
    k1 = t.get("qcut_tally_sliced", 0.0)
    k2 = sum(v for k, v in t.items() if k.startswith("void qcut::reduce") or k.startswith("reduce"))
    step = [v for k, v in t.items() if "literal" not in k and "reduce_partial_kernel<double" not in k]
    f.write(f"K1 share of the default-mode kernels (qcut_tally_sliced / (all but the literal pass)): {k1 / sum(step):.3f}\n")
    print(f"K1 share {k1 / sum(step):.3f}")
PY
