#!/usr/bin/env python
"""Generate tests/golden/frames/ by running the REFERENCE binary (development-time only, see tools/gen_golden.py):
synthetic spacers (+ edge cases) -> createdb -> extractorfs --min-length 9 --orf-start-mode 1 --create-lookup 1 ->
translatenucs, i.e. the query preparation of createsetdb --extractorf-spacer 1 (data/createsetdb.sh:88-113).

Usage:  python tools/gen_golden_frames.py [--ref-bin /tmp/refbuild/src/spacepharer]
Writes sp.fasta.gz (input), aa.txt.gz (key \\t peptide of <db>), hdr.txt.gz (key \\t header of <db>_nucl_orf_h)."""
import argparse
import gzip
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spacepharer_b200 import dbio, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ref-bin", default="/tmp/refbuild/src/spacepharer")
args = ap.parse_args()
out = os.path.join(ROOT, "tests", "golden", "frames")
os.makedirs(out, exist_ok=True)
work = tempfile.mkdtemp(prefix="spb_frames_")

tres, toff, _ = synth.make_targets(n_genomes=20, seed=3)
nt, lens = synth.make_spacers(600, tres, toff, seed=5)
rng = np.random.default_rng(1)
recs = []
for i in range(len(lens)):
    s = "".join("ACGT"[c] for c in nt[i, :lens[i]])
    if i % 37 == 5:        # ambiguous bases
        s = list(s)
        for _ in range(int(rng.integers(1, 3))):
            s[int(rng.integers(0, len(s)))] = "N"
        s = "".join(s)
    recs.append(("sp%d" % i, s))
recs += [("short", "AC"), ("three", "ATG"), ("stopfirst", "TAAATGGCTGCTGCTGCTGCTGCTGCTGCTGCTGCTGCTTGAC"),
         ("stops", "TAATAGTGA" * 5), ("exact27", "GCT" * 9), ("exact26", ("GCT" * 9)[:26]), ("exact29", "GCT" * 9 + "GC"),
         ("endstop", "GCT" * 10 + "TAA"), ("allN", "N" * 40),
         ("long", "".join("ACGT"[c] for c in rng.integers(0, 4, 900))),
         ("longer", "".join("ACGT"[c] for c in rng.integers(0, 4, 5000)))]
fa = os.path.join(work, "sp.fasta")
with open(fa, "w") as f:
    for n, s in recs:
        f.write(">%s\n%s\n" % (n, s))
R = args.ref_bin
db = os.path.join(work, "spdb")
for cmd in ([R, "createdb", fa, db, "--dbtype", "2", "-v", "1"],
            [R, "extractorfs", db, db + "_orf", "--min-length", "9", "--orf-start-mode", "1", "--create-lookup", "1", "--threads", "3", "-v", "1"],
            [R, "translatenucs", db + "_orf", db + "_aa", "--threads", "2", "-v", "1"]):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
with gzip.open(os.path.join(out, "sp.fasta.gz"), "wt") as f:
    f.write(open(fa).read())
keys, seqs = dbio.read_sequence_db(db + "_aa")
with gzip.open(os.path.join(out, "aa.txt.gz"), "wt") as f:
    for k, s in zip(keys, seqs):
        f.write("%d\t%s\n" % (int(k), s.decode() if isinstance(s, bytes) else s))
hdr = dbio.read_entries(db + "_orf_h")
with gzip.open(os.path.join(out, "hdr.txt.gz"), "wt") as f:
    for k in sorted(hdr):
        f.write("%d\t%s\n" % (k, hdr[k].decode().rstrip("\n\0")))
# the input DB keys (createdb may reorder nothing with default --shuffle for nucleotides? record them)
ikeys, iseqs = dbio.read_sequence_db(db)
with gzip.open(os.path.join(out, "input_db.txt.gz"), "wt") as f:
    for k, s in zip(ikeys, iseqs):
        f.write("%d\t%s\n" % (int(k), s.decode() if isinstance(s, bytes) else s))
print("frames:", len(keys))
