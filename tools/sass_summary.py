#!/usr/bin/env python
"""SASS evidence for the hot kernels (runs here, no GPU): registers / shared memory / spills from `cuobjdump -res-usage` and an
opcode histogram + excerpt of the inner loop from `cuobjdump -sass` of the built objects.  Writes profiles/<tag>_sass.txt.
Usage: python tools/sass_summary.py <tag>"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
KERNELS = [("build/prefilter.o", "expand_kernel"), ("build/prefilter.o", "expand_heavy_kernel"), ("build/prefilter.o", "lookup_kernel"), ("build/prefilter.o", "bookkeep_kernel"),
           ("build/align.o", "sw_fast_kernel"), ("build/align.o", "sw_exhaustive_kernel"), ("build/align.o", "sw_traceback_kernel"),
           ("build/align.o", "sw_reverse_kernel")]
INTEREST = ("VIADDMNMX", "VIMNMX", "VIADD", "PRMT", "LDG", "LDS", "STG", "STS", "ATOMS", "ATOMG", "RED", "SHFL", "REDUX", "VOTE", "POPC",
            "IMAD", "LOP3", "BAR", "DADD", "DMUL", "DFMA", "F2F", "LDC", "ISETP", "BRA")
out = ["# SASS summary of the hot kernels (cuobjdump, sm_100a objects of this commit); tag %s" % tag, ""]
res = {}
for obj in sorted({o for o, _ in KERNELS}):
    txt = subprocess.run(["cuobjdump", "-res-usage", os.path.join(ROOT, obj)], capture_output=True, text=True).stdout
    for m in re.finditer(r"Function (\S+):\n\s+(REG:\d+.*)", txt):
        res[(obj, m.group(1))] = m.group(2)
for obj, kern in KERNELS:
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, obj)], capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : ", sass)
    for part in parts[1:]:
        name = part.split("\n", 1)[0].strip()
        if kern not in name:
            continue
        ops = collections.Counter()
        lines = []
        for l in part.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if m:
                ops[m.group(1).split(".")[0]] += 1
                lines.append(l.rstrip())
        total = sum(ops.values())
        out.append("## %s  (%s)" % (name, obj))
        out.append("resources: " + res.get((obj, name), "?"))
        out.append("instructions: %d static; " % total + ", ".join("%s %d" % (k, ops[k]) for k in INTEREST if ops.get(k)))
        dpx = [l for l in lines if "VIADDMNMX" in l or "VIMNMX" in l][:6]
        if dpx:
            out.append("DPX / SIMD-video excerpt:")
            out += ["    " + re.sub(r"\s+/\*[0-9a-f]{16}\*/", "", l.strip()) for l in dpx]
        ldg = [l for l in lines if re.search(r"\bLDG|\bLD\.E|LDC\.64", l)][:4]
        if ldg and not dpx:
            out.append("global-load excerpt:")
            out += ["    " + re.sub(r"\s+/\*[0-9a-f]{16}\*/", "", l.strip()) for l in ldg]
        out.append("")
path = os.path.join(ROOT, "profiles", "%s_sass.txt" % tag)
open(path, "w").write("\n".join(out) + "\n")
print("wrote", path)
