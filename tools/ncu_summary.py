#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the per-launch figures the bench line and DESIGN.md quote:
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/<name>.json "<command that was profiled>"
Reads the report with `ncu -i <rep> --page raw --csv` (no GPU needed)."""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
rep, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
launches = []
for r in data:
    rec = {"kernel": r[ki][:100]}
    for name in KEEP:
        if name in hdr:
            i = hdr.index(name)
            rec[name] = f"{r[i]} {units[i]}".strip()
    launches.append(rec)
json.dump({"source": f"ncu --set full --clock-control none, read with ncu -i {rep} --page raw --csv", "command": cmd, "launches": launches},
          open(dst, "w"), indent=1)
for l in launches:
    print(l["kernel"][:60], l.get("gpu__time_duration.sum"), l.get("dram__bytes_read.sum"), l.get("dram__bytes_write.sum"),
          l.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))
