#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/build.log 2>&1 || { tail -20 $OUT/build.log; exit 1; }
CMD="python bench.py --workload ${WL:-cfg4_heavy_hex} --scale ${SCALE:-1.0} --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > $OUT/launches_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 40 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r)>10]
hdr=rows[0]; 
for r in rows[1:]:
    d=dict(zip(hdr,r))
    if d["Metric Name"]=="gpu__time_duration.sum": print(d["ID"], d["Kernel Name"][:60], d["Grid Size"], d["Block Size"], d["Metric Value"], d["Metric Unit"])
PY
