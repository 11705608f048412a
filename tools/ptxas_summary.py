#!/usr/bin/env python
"""Summarise `-Xptxas -v` logs: registers, spills, shared memory per kernel."""
import re, subprocess, glob, os
here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spectralmc_b200", "csrc")
for f in sorted(glob.glob(os.path.join(here, "*.ptxas.log"))):
    txt = open(f).read()
    for m in re.finditer(r"Compiling entry function '(\S+)'.*\n.*\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*Used (\d+) registers(.*)", txt):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        print(f"regs={m.group(5):>3} stack={m.group(2):>4} spill={m.group(3)}/{m.group(4)} {m.group(6).strip(', '):28s} {name}")
