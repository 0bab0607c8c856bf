"""Summarise an .ncu-rep (raw page) into the handful of metrics DESIGN.md / profiles/ quote.
usage: python scripts/ncu_summary.py report.ncu-rep [kernel-regex]"""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum.per_cycle_active", "sm__inst_executed.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sass__inst_executed_shared_loads",
    "sass__inst_executed_global_loads", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]
name_i = hdr.index("Kernel Name")
for r in rows[2:]:
    if pat and not pat.search(r[name_i]):
        continue
    print("==", r[name_i][:100])
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    for k in KEYS:
        if k in d:
            print(f"  {k:75s} {d[k]:>18s} {u[k]}")
    stalls = [(h, float(d[h])) for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and d[h]]
    stalls.sort(key=lambda x: -x[1])
    for h, v in stalls[:8]:
        print(f"  stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {v:.3f}")
