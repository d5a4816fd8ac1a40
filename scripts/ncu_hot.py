#!/usr/bin/env python
"""Hot spots of a kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view): executed-instruction
share by opcode, and the SASS ranges with the most stall samples.

  python scripts/ncu_hot.py gpurun_out/C3_noloss.ncu-rep
"""
import collections
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = [r for r in rows[2:] if len(r) > iex]
tot_ex = sum(int(r[iex]) for r in body)
tot_smp = sum(int(r[ismp]) for r in body)
print("kernel:", rows[0][1][:100])
print("warp-instructions executed: %d   stall samples: %d" % (tot_ex, tot_smp))
by_op = collections.Counter()
smp_op = collections.Counter()
for r in body:
    toks = r[isrc].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    by_op[op] += int(r[iex])
    smp_op[op] += int(r[ismp])
print("\nopcode        exec%   samples%")
for op, n in by_op.most_common(28):
    print("%-12s %6.2f   %6.2f" % (op, 100.0 * n / tot_ex, 100.0 * smp_op[op] / max(tot_smp, 1)))
print("\ntop instructions by stall samples:")
order = sorted(range(len(body)), key=lambda i: -int(body[i][ismp]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i in order:
    r = body[i]
    print("%5d  smp %5.2f%%  exec %5.2f%%  %s" % (i, 100.0 * int(r[ismp]) / max(tot_smp, 1), 100.0 * int(r[iex]) / tot_ex,
                                                r[isrc].strip()[:90]))
