#!/usr/bin/env python3
"""Turn an ncu report (.ncu-rep) into the small tracked summary under profiles/.
usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/name   -> name.raw.csv (selected metrics), name.md"""
import csv
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum",
    "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmalite.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__sass_thread_inst_executed_op_integer_pred_on.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(hdr) if h in ("ID", "Kernel Name", "Grid Size", "Block Size") or h in KEEP or (h.startswith(STALL) and h.endswith("_per_issue_active.ratio"))]
    with open(out + ".raw.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in cols])
        w.writerow([units[i] for i in cols])
        for r in rows[2:]:
            w.writerow([r[i] for i in cols])
    with open(out + ".md", "w") as f:
        f.write(f"# ncu summary of `{rep.split('/')[-1]}` (ncu --set full --clock-control none --import-source on)\n\n")
        for r in rows[2:]:
            f.write(f"## launch {r[hdr.index('ID')]}: `{r[hdr.index('Kernel Name')][:80]}` grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}\n\n")
            f.write("| metric | value | unit |\n|---|---:|---|\n")
            for i in cols:
                if hdr[i] in ("ID", "Kernel Name", "Grid Size", "Block Size"):
                    continue
                f.write(f"| {hdr[i]} | {r[i]} | {units[i]} |\n")
            f.write("\n")


if __name__ == "__main__":
    main()
