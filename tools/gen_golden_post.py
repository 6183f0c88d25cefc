#!/usr/bin/env python
"""Inputs of the reference's post-processing (data/predictmatch.sh:127-168) for the example run (BASELINE config 1), so that the
final result.tsv can be re-derived from OUR `aggregate` / `match` DBs by the reference's own downstream modules on the GPU box
(tests/test_gpu_cli.py::test_final_tsv_identical_to_reference).  Development time only: runs the reference binary
(oracle/_ref/spacepharer) on /root/reference/examples and packs the DB families the six post-processing modules read:
  tests/golden/config1_post/targets.tar.gz   createsetdb output of the three example phages (targets, targets_nucl*, headers, lookups)
  tests/golden/config1_post/qsetdb.tar.gz    createsetdb --extractorf-spacer 1 output of the two CRISPR arrays (qsetdb, qsetdb_nucl*, ...)
and checks that the run reproduces tests/golden/config1/result.tsv.
Usage: python tools/gen_golden_post.py [--ref-src /root/reference]"""
import argparse
import glob
import os
import shutil
import subprocess
import sys
import tarfile
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refbin  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ref-src", default="/root/reference")
a = ap.parse_args()
assert refbin.available(), "build the reference first: make -C oracle ref"
B = refbin.REF_BIN
ex = os.path.join(a.ref_src, "examples")
w = tempfile.mkdtemp(prefix="spb_post_")
run = lambda cmd: subprocess.check_call(cmd, shell=True, cwd=w, stdout=subprocess.DEVNULL)
gca = " ".join(sorted(glob.glob(os.path.join(ex, "GCA*.fna.gz"))))
fas = " ".join(sorted(glob.glob(os.path.join(ex, "*.fas"))))
run("%s createsetdb %s targets tmp --threads 4 -v 1" % (B, gca))
run("%s createsetdb %s targets_rev tmp --reverse-fragments 1 --threads 4 -v 1" % (B, gca))
run("%s easy-predict %s targets result.tsv tmp --threads 4 --remove-tmp-files 0 -v 1" % (B, fas))
gold = os.path.join(ROOT, "tests", "golden", "config1", "result.tsv")
assert open(os.path.join(w, "result.tsv"), "rb").read() == open(gold, "rb").read(), "this run does not reproduce the committed result.tsv"
out = os.path.join(ROOT, "tests", "golden", "config1_post")
os.makedirs(out, exist_ok=True)
qdir = os.path.dirname(glob.glob(os.path.join(w, "tmp", "*", "qsetdb.index"))[0])
for name, d, pat, skip in (("targets", w, "targets*", "targets_rev"), ("qsetdb", qdir, "qsetdb*", None)):
    with tarfile.open(os.path.join(out, name + ".tar.gz"), "w:gz", dereference=True) as t:
        for f in sorted(glob.glob(os.path.join(d, pat))):
            b = os.path.basename(f)
            if skip and b.startswith(skip):
                continue
            t.add(f, arcname=b)
shutil.rmtree(w)
print("written", out)
