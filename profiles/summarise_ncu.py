"""Turns an .ncu-rep (ncu --set full) into the short text summary kept under profiles/.
usage: python profiles/summarise_ncu.py report.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_<kernel>.txt"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg",
]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
        print(d.get("Kernel Name", "?")[:100])
        for k in KEYS:
            if k in d and d[k] != "":
                print(f"    {k} = {d[k]} {u[k]}")
        st = [(float(v), h) for h, v in d.items()
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and v]
        for v, h in sorted(st, reverse=True)[:6]:
            name = h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
            print(f"    stall[{name}] = {v:.3f} warps per issue slot")
