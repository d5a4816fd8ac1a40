#!/usr/bin/env python
"""Per-source-line share of executed warp instructions and of warp-stall samples from an ncu report captured with
`--set full --import-source on` (the library is built with -lineinfo), as a markdown table.

  python scripts/ncu_source_stalls.py gpurun_out/step_c2.ncu-rep [top_n] > profiles/<name>.md
"""
import csv, io, subprocess, sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur, hdr, kernel, agg, worst = None, None, None, {}, {}
tot_i = tot_s = 0
line_key = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split('/')[-1]
    elif r[0] == "Function Name":
        kernel = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 3:
        d = dict(zip(hdr, r))
        if r[2] == '-' and r[0].strip().isdigit():
            try:
                inst, samp = int(d["Instructions Executed"]), int(d["# Samples"])
            except ValueError:
                continue
            line_key = (cur, int(r[0]))
            agg[line_key] = (inst, samp, r[1].strip())
            tot_i += inst
            tot_s += samp
        elif line_key and r[2] not in ('-', '... ...') and d.get("# Samples", "").isdigit():
            s = int(d["# Samples"])
            if s > worst.get(line_key, (0, "", {}))[0]:
                stalls = {k[6:]: int(v) for k, v in d.items()
                          if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
                worst[line_key] = (s, r[3].strip(), stalls)
print("Kernel: `%s`\n" % kernel)
print("%d warp instructions executed, %d warp-stall samples.\n" % (tot_i, tot_s))
print("| source line | % of instructions | % of stall samples | most-stalled SASS instruction (samples: top reasons) |")
print("|---|---|---|---|")
for (f, l), (inst, samp, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top_n]:
    w = worst.get((f, l))
    ws = ""
    if w:
        top = sorted(w[2].items(), key=lambda kv: -kv[1])[:2]
        ws = "`%s` (%d: %s)" % (w[1][:48], w[0], ", ".join("%s %d" % kv for kv in top))
    print("| %s:%d `%s` | %.1f | %.1f | %s |" % (f, l, src[:70].replace('|', '\\|'), 100.0 * inst / tot_i,
                                              100.0 * samp / tot_s, ws))
