"""python tools/ncu_summary.py <report.ncu-rep> <out.csv>: the per-launch metrics the roofline discussion uses, one row per
captured launch (read with `ncu -i ... --page raw --csv`, as /opt/skills/guides/B200_PROFILING.md prescribes)."""
import csv, subprocess, sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "sm__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
cols = [i for i, h in enumerate(hdr) if h in ("ID", "Kernel Name") or h in KEEP]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in cols])
    for r in rows[2:]:
        w.writerow([r[i][:110] for i in cols])
print(open(sys.argv[2]).read()[:3000])
