"""DRAM traffic of the K1 launches captured in an .ncu-rep -> profiles/<round>_k1_traffic.json (read by bench.py).

    python scripts/ncu_traffic.py gpurun_out/prof_k1.ncu-rep profiles/r01_k1_traffic.json cfg4_heavy_hex 1.0
"""
import csv
import json
import subprocess
import sys

rep, out, workload, scale = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
units = rows[1]
total, launches = 0.0, []
for r in rows[2:]:
    rd = float(r[hdr.index("dram__bytes_read.sum")]) * unit[units[hdr.index("dram__bytes_read.sum")]]
    wr = float(r[hdr.index("dram__bytes_write.sum")]) * unit[units[hdr.index("dram__bytes_write.sum")]]
    dur = float(r[hdr.index("gpu__time_duration.sum")])
    launches.append({"kernel": r[hdr.index("Kernel Name")], "dram_read_bytes": rd, "dram_write_bytes": wr,
                     "duration": dur, "duration_unit": units[hdr.index("gpu__time_duration.sum")]})
    total += rd + wr
json.dump({"workload": workload, "scale": scale, "k1_kernel_launches_per_step": len(launches),
           "traffic_bytes_per_k1_launch": total, "launches": launches,
           "source": f"ncu --set full capture {rep} (dram__bytes_read.sum + dram__bytes_write.sum)"},
          open(out, "w"), indent=1)
print(out, total)
