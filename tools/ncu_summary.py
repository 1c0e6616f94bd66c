"""Summarise one kernel of an .ncu-rep (ncu --set full) as the markdown table profiles/ keeps:
    python tools/ncu_summary.py gpurun_out/prof_ell_cfg2_f64_r01_final.ncu-rep > /tmp/summary.md
Reads the report with `ncu -i ... --page raw --csv` (works without a GPU)."""
import csv
import io
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}
print("| metric | unit | value |\n|---|---|---|")
for k in KEYS:
    if k in col:
        print("| %s | %s | %s |" % (k, units[col[k]], vals[col[k]]))
stalls = []
for h, i in col.items():
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        try:
            stalls.append((float(vals[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        except ValueError:
            pass
print("\nWarp stall reasons (cycles per issue-active cycle):\n")
for v, name in sorted(stalls, reverse=True)[:8]:
    print("- %s: %.2f" % (name, v))
