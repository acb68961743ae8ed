"""Text summary (the format of profiles/*_ncu.txt) of every kernel launch in an .ncu-rep captured with --set full.

    python tools/ncu_text.py gpurun_out/x.ncu-rep profiles/x_ncu.txt ["header line"]

Per launch: duration, launch shape, instruction counts, pipe utilisation (XU = MUFU pipe, FMA, ALU, FP64, LSU), issue
activity, eligible / active warps, DRAM bytes and the warp-state breakdown."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    head = sys.argv[3] if len(sys.argv) > 3 else f"ncu --set full --clock-control none ({rep})"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    lines = [f"# {head}"]
    for n, r in enumerate(launches):
        name = r[hdr.index("Kernel Name")]
        lines.append(f"## launch {n}: {name[:110]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"{k:78s} {r[i]:>16s} {units[i]}")
        stalls = [(h[len(STALL):].replace("_per_issue_active.ratio", ""), float(r[i].replace(",", "")))
                  for i, h in enumerate(hdr) if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
        lines.append("# warp states per issue-active cycle (" + STALL + "*_per_issue_active.ratio)")
        for nm, v in sorted(stalls, key=lambda x: -x[1]):
            if v >= 0.05:
                lines.append(f"  {nm:32s} {v:6.3f}")
    open(out, "w").write("\n".join(lines) + "\n")
    print(f"{len(launches)} launches -> {out}")


if __name__ == "__main__":
    main()
