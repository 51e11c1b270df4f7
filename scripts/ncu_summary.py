"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote.

    python scripts/ncu_summary.py gpurun_out/prof_k1.ncu-rep
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_requests_srcunit_tex_op_read.sum",
    "memory_l2_theoretical_sectors_global", "memory_l2_theoretical_sectors_global_ideal", "dram__sectors_read.sum",
]
for r in rows[2:]:
    print("=== kernel", r[hdr.index("Kernel Name")], "launch id", r[0])
    vals = {}
    for i, h in enumerate(hdr):
        if h in KEYS:
            print(f"  {h:75s} {units[i]:12s} {r[i]}")
        if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                vals[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = float(r[i])
            except ValueError:
                pass
    def num(name):
        try:
            return float(r[hdr.index(name)].replace(",", ""))
        except (ValueError, IndexError):
            return float("nan")

    # L2 / sector efficiency of the shot-byte stream (north star): sectors per global-load request (a fully used warp-wide
    # 128-bit load is 16 sectors of 32 B), bytes the L2 delivered to the SMs per byte read from DRAM, and the
    # requested-vs-ideal sector ratio ncu derives for global loads (1.0 = no partially used sectors)
    sec, req = num("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"), num("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")
    l2s, drs = num("lts__t_sectors_srcunit_tex_op_read.sum"), num("dram__sectors_read.sum")
    the, ideal = num("memory_l2_theoretical_sectors_global"), num("memory_l2_theoretical_sectors_global_ideal")
    print(f"  derived: global-load sectors per request {sec / req if req else float('nan'):.2f}; L2->SM read sectors / DRAM read sectors "
          f"{l2s / drs if drs else float('nan'):.3f}; requested / ideal L2 sectors of global accesses {the / ideal if ideal else float('nan'):.3f}")
    top = sorted(vals.items(), key=lambda kv: -kv[1])[:8]
    print("  stalls (warps per issue-active cycle):", ", ".join(f"{k} {v:.2f}" for k, v in top))
