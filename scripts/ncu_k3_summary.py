"""Summarise an `ncu --set full` report of scripts/profile_k3.py (K3a / K3b / reduce) into the profiles/ text format."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum"]
ik = hdr.index("Kernel Name")
for r in rows[2:]:
    print(f"== {r[ik][:60]}  (ncu --set full --clock-control none, mb=32768, scripts/profile_k3.py)")
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:75s} {r[i]} {units[i]}")
