#!/usr/bin/env python
"""Summarise gpurun_out/launches.csv + gpurun_out/prof.ncu-rep into profiles/ (tracked)."""
import collections
import csv
import json
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(root, "profiles")
os.makedirs(out_dir, exist_ok=True)
go = os.path.join(root, "gpurun_out")

rows = [r for r in csv.reader(open(os.path.join(go, "launches.csv"))) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    except ValueError:
        pass
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(out_dir, "%s_launches.md" % tag), "w") as f:
    f.write("# %s: ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n" % tag)
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv python bench.py "
            "--steps 20 --warmup 3 --no-e2e --no-cpu` (cold-cache, serialised: compare shares).\n\n")
    f.write("| kernel | launches | mean us | share |\n|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write("| `%s` | %d | %.1f | %.3f |\n" % (k[:110], len(v), sum(v) / len(v) / 1e3, sum(v) / tot))
    step = [k for k in agg if "elbo_step" in k]
    upd = [k for k in agg if "event_update" in k]
    if step and upd:
        s, u = sum(agg[step[0]]), sum(agg[upd[0]])
        f.write("\nWithin an optimisation step (elbo_step + event_update): fused step kernel share = %.3f\n" % (s / (s + u)))

raw = subprocess.run(["ncu", "-i", os.path.join(go, "prof.ncu-rep"), "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
name_i = h.index("Kernel Name")
with open(os.path.join(out_dir, "%s_step_kernel.md" % tag), "w") as f:
    f.write("# %s: ncu --set full of the fused step kernel\n\n" % tag)
    f.write("Command: `ncu --set full --clock-control none --import-source on -k regex:elbo_step -s 4 -c 2 python "
            "bench.py --steps 6 --warmup 3 --no-e2e --no-cpu` (C2: 5000 x 5000, M=2, S=3).\n\n")
    f.write("Kernel: `%s`\n\n| metric | unit | launch 1 | launch 2 |\n|---|---|---|---|\n" % rr[2][name_i])
    vals = {}
    for w in want:
        if w in h:
            i = h.index(w)
            col = [r[i] for r in rr[1:]]
            vals[w] = col
            f.write("| %s | %s | %s |\n" % (w, col[0], " | ".join(col[1:3])))

    def gb(name):
        unit, v = vals[name][0], float(vals[name][1])
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}[unit]
    traffic = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
    f.write("\nDRAM traffic per launch = %.4f GB (read %.4f + write %.4f); algorithmic bytes per launch = 2.7000 GB "
            "(25e6 cell-events x (12 + 48 x 2) B).\n" % (traffic / 1e9, gb("dram__bytes_read.sum") / 1e9,
                                                         gb("dram__bytes_write.sum") / 1e9))
json.dump({"dram_bytes_per_launch": traffic, "source": "profiles/%s_step_kernel.md" % tag,
           "kernel": "elbo_step_kernel<1,0,false,false>", "workload": "C2 5000x5000 M=2 S=3"},
          open(os.path.join(out_dir, "traffic.json"), "w"), indent=1)
print(open(os.path.join(out_dir, "%s_launches.md" % tag)).read())
print(open(os.path.join(out_dir, "%s_step_kernel.md" % tag)).read())
