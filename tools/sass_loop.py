#!/usr/bin/env python
"""Print the instruction mix of the innermost loops of one kernel in the built library.

usage: python tools/sass_loop.py <mangled-name-prefix> [--dump]
"""
import collections, re, subprocess, sys, os
lib = os.environ.get("SPECTRALMC_B200_LIB") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spectralmc_b200", "lib", "libspectralmc_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    if not part.startswith(sys.argv[1]):
        continue
    ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4})\*/\s+(.*?);", part)]
    print(part.split("\n")[0], len(ins), "instructions")
    loops = []
    for a, t in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            loops.append((int(m.group(1), 16), a))
    for lo, hi in loops:
        if any(l2 > lo and h2 < hi for l2, h2 in loops):
            continue  # not innermost
        body = [t for a, t in ins if lo <= a <= hi]
        mix = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", b).split()[0] for b in body)
        print(f"loop {lo:#x}..{hi:#x}: {len(body)} instr:", ", ".join(f"{v} {k}" for k, v in mix.most_common()))
        if "--dump" in sys.argv:
            print("\n".join("    " + b for b in body))
