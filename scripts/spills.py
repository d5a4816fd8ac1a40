#!/usr/bin/env python
"""Registers / spill bytes of every elbo_step_kernel instantiation from a `build.py -v --force` log."""
import re
import sys

txt = open(sys.argv[1]).read().split('\n')
want = set(sys.argv[2:])
cur = None
for i, l in enumerate(txt):
    m = re.search(r"elbo_step_kernelILi(\d+)ELi(\d+)ELb([01])ELb([01])", l)
    if m and 'Compiling' in l:
        cur = m.groups()
    elif cur and 'spill' in l:
        sp = re.findall(r"(\d+) bytes", l)
        reg = re.search(r"Used (\d+) registers", txt[i + 1])
        key = "KC%s_KG%s_%s_%s" % (cur[0], cur[1], "cell" if cur[2] == '1' else "gene", "loss" if cur[3] == '1' else "noloss")
        if not want or key in want:
            print(key, "stack/st/ld bytes", sp, "regs", reg.group(1) if reg else None)
        cur = None
