#!/usr/bin/env python
"""Instruction mix (warp-level executed counts by opcode) of one kernel from an .ncu-rep source page.
    python tools/ncu_instmix.py rep.ncu-rep regex:k_name [top]"""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
si, ie, ss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
ops, samp = collections.Counter(), collections.Counter()
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) < len(hdr):
        continue
    try:
        cnt, st = int(r[ie]), int(r[ss])
    except ValueError:
        continue
    t = r[si].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += cnt
    samp[op] += st
tot = sum(ops.values())
print(f"total warp instructions {tot}")
for op, c in ops.most_common(top):
    print(f"{op:10s} {c:10d} {c / tot:.3f}  stall-samples {samp[op]}")
