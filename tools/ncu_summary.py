"""Condense an .ncu-rep of the implicit kernel into the small JSON bench.py reads for its roofline block.

    python tools/ncu_summary.py gpurun_out/implicit_r01c.ncu-rep profiles/implicit_kernel_ncu_summary.json

The issued MUFU (XU pipe) instruction count is deterministic for the seeded bench workload, so it is taken from the
profile once; bench.py divides it by the kernel time it measures live.
"""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, launches = rows[0], rows[2:]


def col(r, name):
    return float(r[hdr.index(name)].replace(",", ""))


per = []
for r in launches:
    cyc = col(r, "sm__cycles_elapsed.avg")
    sms = 148
    xu_pct = col(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed")
    per.append({
        "kernel": r[hdr.index("Kernel Name")][:60],
        "duration_us": col(r, "gpu__time_duration.sum"),
        "sm_cycles_elapsed": cyc,
        "xu_pct_of_peak_elapsed": xu_pct,
        "xu_pct_of_peak_active": col(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "xu_warp_inst": xu_pct / 100.0 * 0.5 * sms * cyc,          # peak = 16 thread-ops = 0.5 warp-inst / clk / SM
        "warp_inst_executed": col(r, "smsp__inst_executed.sum"),
        "issue_active_pct": col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "fma_pipe_pct_active": col(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        "alu_pipe_pct_active": col(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": col(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": col(r, "launch__registers_per_thread"),
        "grid": col(r, "launch__grid_size"), "block": col(r, "launch__block_size"),
        "dram_bytes": col(r, "dram__bytes_read.sum") * {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}[rows[1][hdr.index("dram__bytes_read.sum")]]
                      + col(r, "dram__bytes_write.sum") * {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}[rows[1][hdr.index("dram__bytes_write.sum")]],
    })
n = len(per)
summary = {"source": rep, "launches": n,
           "xu_warp_inst_per_launch": sum(p["xu_warp_inst"] for p in per) / n,
           "warp_inst_per_launch": sum(p["warp_inst_executed"] for p in per) / n,
           "dram_bytes_per_launch": sum(p["dram_bytes"] for p in per) / n,
           "duration_us_under_ncu": sum(p["duration_us"] for p in per) / n,
           "per_launch": per}
json.dump(summary, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in summary.items() if k != "per_launch"}))
