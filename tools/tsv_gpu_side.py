#!/usr/bin/env python
"""GPU side of the final-TSV parity check (tools/tsv_parity_check.sh): run `spb200 predictmatch` (default pipeline) on the
reference's example DBs (tests/golden/config1) and leave the `aggregate` and `match` DBs under gpurun_out/tsv/."""
import gzip, os, shutil, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C1 = os.path.join(ROOT, "tests", "golden", "config1")
w = tempfile.mkdtemp()
for f in os.listdir(C1):
    src = os.path.join(C1, f)
    if f.endswith("_nucl_orf.gz"):
        with gzip.open(src, "rb") as g, open(os.path.join(w, f[:-3]), "wb") as o:
            shutil.copyfileobj(g, o)
    elif not f.endswith(".txt.gz") and not f.endswith(".tsv"):
        shutil.copyfile(src, os.path.join(w, f))
out = os.path.join(ROOT, "gpurun_out", "tsv")
shutil.rmtree(out, ignore_errors=True)
os.makedirs(out)
r = subprocess.run([os.path.join(ROOT, "spacepharer_b200", "lib", "spb200"), "predictmatch", w + "/qsetdb", w + "/targets", w + "/targets_rev",
                    out + "/match", out, "--threads", "1", "--keep-intermediate", "1"], capture_output=True, text=True)
print(r.stdout[-400:], r.stderr[-400:])
sys.exit(r.returncode)
