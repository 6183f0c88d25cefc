#!/usr/bin/env python
"""Per-source-line stall samples / instruction counts of one kernel from an ncu report (needs -lineinfo).
Usage: python tools/ncu_lines.py <report.ncu-rep> <object.o> <mangled kernel name> <source.cu> [launch-skip] [top]"""
import csv, io, re, subprocess, sys, tempfile, os
rep, obj, kern, src = sys.argv[1:5]
skip = sys.argv[5] if len(sys.argv) > 5 else "0"
top = int(sys.argv[6]) if len(sys.argv) > 6 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text." + kern + ":")][0]
cur, amap = None, {}
for l in dis[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        amap[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][:2])
hdr = rows[1]
ia, isamp, iins = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base, per, ts, ti = None, {}, 0, 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    except ValueError:
        continue
    if base is None:
        base = a
    s, n = int(r[isamp] or 0), int(r[iins] or 0)
    ln = amap.get(a - base)
    p = per.setdefault(ln, [0, 0])
    p[0] += s; p[1] += n; ts += s; ti += n
text = open(src).read().split("\n")
print("samples", ts, "warp instructions", ti)
for ln, (s, n) in sorted(per.items(), key=lambda x: -x[1][1 if os.environ.get("BY_INS") else 0])[:top]:
    t = text[ln[1] - 1].strip()[:100] if ln and ln[0] == os.path.basename(src) else str(ln)
    print("%6.2f%% smp %6.2f%% ins  %-6s %s" % (100 * s / max(ts, 1), 100 * n / max(ti, 1), ln[1] if ln else "?", t))
# optional phase totals: RANGES="name:lo-hi,name:lo-hi" (source lines of <source.cu>; other files are listed under their own name)
if os.environ.get("RANGES"):
    rng = []
    for part in os.environ["RANGES"].split(","):
        nm, r = part.split(":")
        lo, hi = r.split("-")
        rng.append((nm, int(lo), int(hi)))
    tot = {}
    for ln, (s, n) in per.items():
        nm = "?"
        if ln and ln[0] == os.path.basename(src):
            nm = next((a for a, lo, hi in rng if lo <= ln[1] <= hi), "other")
        elif ln:
            nm = ln[0]
        p = tot.setdefault(nm, [0, 0])
        p[0] += s; p[1] += n
    print("--- phases")
    for nm, (s, n) in sorted(tot.items(), key=lambda x: -x[1][1]):
        print("%6.2f%% smp %6.2f%% ins  %s" % (100 * s / max(ts, 1), 100 * n / max(ti, 1), nm))
