#!/usr/bin/env python
"""One line per profiled launch of an `ncu --set full` report: duration, DRAM bytes, and the utilisation metrics the rooflines
are argued from.
    python tools/ncu_summary.py gpurun_out/r02_kernels.ncu-rep profiles/r02_ncu_kernels_summary.csv"""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
cols = [c for c in cols if c in ix]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch", "kernel"] + [f"{c} [{units[ix[c]]}]" for c in cols])
    for i, r in enumerate(data):
        name = r[ix["Kernel Name"]]
        if "at::" in name:
            continue
        w.writerow([i, name.split("(")[0].replace("void ", "")] + [r[ix[c]] for c in cols])
print(open(out).read())
