"""Markdown summary of `ncu --set full` captures (.ncu-rep) for profiles/: per captured launch the metrics the roofline
argument uses.  Usage: python scripts/ncu_summary.py title rep1.ncu-rep [rep2 ...] > profiles/ncu_rN_summary.md"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
STALLS = ["barrier", "wait", "long_scoreboard", "short_scoreboard", "mio_throttle", "math_pipe_throttle", "not_selected", "no_instruction",
          "branch_resolving", "lg_throttle", "dispatch_stall", "selected"]
print(f"# {sys.argv[1]}\n")
for rep in sys.argv[2:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for row in rows[2:]:
        name = row[h.index("Kernel Name")]
        print(f"## {name[:110]}\n\nsource: `{rep.split('/')[-1]}` (launch id {row[0]})\n\n| metric | value |\n|---|---|")
        for k in KEYS:
            if k in h:
                i = h.index(k)
                print(f"| {k} | {row[i]} {u[i]} |")
        st = []
        for s in STALLS:
            k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if k in h:
                try:
                    st.append((float(row[h.index(k)]), s))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("| warp-stall mix (warps stalled per issued instruction) | " + ", ".join(f"{s} {v:.2f}" for v, s in st[:8]) + " |\n")
