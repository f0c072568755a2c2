#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xyz_summary.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    lines = [f"# ncu --set full --clock-control none summary of {rep}", "# one column per captured launch", ""]
    name_i = hdr.index("Kernel Name")
    lines.append(f"{'Kernel Name':78s} " + " | ".join(r[name_i] for r in data))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"{k:78s} {units[i]:16s} " + " | ".join(r[i] for r in data))
    stalls = []
    for i, h in enumerate(hdr):
        if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                stalls.append((float(data[0][i]), h))
            except ValueError:
                pass
    lines.append("")
    lines.append("# warps stalled per issued instruction, by reason (first launch)")
    for v, h in sorted(stalls, reverse=True)[:12]:
        lines.append(f"{v:8.3f}  {h}")
    # algorithmic vs measured traffic
    try:
        rd = float(data[0][hdr.index('dram__bytes_read.sum')]); wr = float(data[0][hdr.index('dram__bytes_write.sum')])
        u = units[hdr.index('dram__bytes_read.sum')]
        lines.append("")
        lines.append(f"# DRAM traffic per launch: read {rd} + write {wr} = {rd + wr:.1f} {u}")
    except Exception:
        pass
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
