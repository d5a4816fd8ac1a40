#!/usr/bin/env python
"""Digest of one or more .ncu-rep files (ncu --set full): key throughput, occupancy and
stall metrics per captured launch, as a markdown table.

  python scripts/ncu_digest.py gpurun_out/c3.ncu-rep [more.ncu-rep ...] > profiles/xyz.md
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]
STALL = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "barrier", "mio_throttle", "math_pipe_throttle",
          "lg_throttle", "not_selected", "dispatch_stall", "branch_resolving", "no_instruction", "membar",
          "drain", "imc_miss", "tex_throttle", "sleeping", "selected"]


def digest(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units, rows = rr[0], rr[1], rr[2:]
    ni = h.index("Kernel Name")
    print("## %s\n" % path)
    for r in rows:
        print("Kernel: `%s`\n" % r[ni][:160])
        print("| metric | unit | value |\n|---|---|---|")
        for w in WANT + [STALL % s for s in STALLS]:
            if w in h:
                i = h.index(w)
                print("| %s | %s | %s |" % (w, units[i], r[i]))
        print()


if __name__ == "__main__":
    for p in sys.argv[1:]:
        digest(p)
