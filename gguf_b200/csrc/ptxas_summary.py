#!/usr/bin/env python
"""Summarise build/*.ptxas.log: registers, spills, smem per kernel (used by DESIGN.md / profiles)."""
import glob, os, re, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
for f in sorted(glob.glob(os.path.join(here, "build", "*.ptxas.log"))):
    txt = open(f).read()
    blocks = re.findall(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                        r"ptxas info\s+: Used (\d+) registers(?:, used \d+ barriers)?(?:, (\d+) bytes smem)?", txt)
    if not blocks:
        continue
    names = subprocess.run(["c++filt"] + [b[0] for b in blocks], capture_output=True, text=True).stdout.strip().split("\n")
    print(os.path.basename(f))
    for n, b in zip(names, blocks):
        n = re.sub(r"\(.*", "", n).replace("void ggq::", "")
        print(f"  {n:<60s} regs={b[4]:>3s} stack={b[1]} spill={b[2]}/{b[3]} smem={b[5] or 0}")
