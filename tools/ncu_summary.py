"""Condense an `ncu --set full` report into the handful of metrics the roofline discussion uses.
usage: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r01_x_ncu_summary.csv"""
import csv, subprocess, sys

KEYS = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct',
        'lts__throughput.avg.pct', 'l1tex__throughput.avg.pct', 'sm__throughput.avg.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct',
        'sm__inst_executed_pipe_tensor', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'launch__waves_per_multiprocessor', 'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct', 'gpc__cycles_elapsed.avg.per_second')

def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = [i for i, h in enumerate(hdr) if h in ('ID', 'Kernel Name', 'Grid Size', 'Block Size') or any(k in h for k in KEYS)]
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in keep])
        w.writerow([units[i] for i in keep])
        for r in rows[2:]:
            w.writerow([r[i] for i in keep])
    print(f"{rep}: {len(rows) - 2} kernels, {len(keep)} columns -> {out}")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
