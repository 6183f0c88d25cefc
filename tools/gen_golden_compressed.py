#!/usr/bin/env python
"""Compressed-DB fixtures from the REFERENCE binary (oracle/_ref/spacepharer); development time only.

  tests/golden/compressed/{targets,qsetdb,pref_0,prot_result}   the DBs of tests/golden/config1 written by the reference's
                                                                `compress` module (DBWriter in WRITER_COMPRESSED_MODE:
                                                                zstd frames for entries of 60 bytes and more, short entries stored;
                                                                DBWriter.cpp:254-397) -- what `--compressed 1` leaves on disk
Usage: python tools/gen_golden_compressed.py
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refbin  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    assert refbin.available(), "build the reference first: make -C oracle ref"
    out = os.path.join(GOLD, "compressed")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    for db in ("targets", "qsetdb", "pref_0", "prot_result"):
        subprocess.check_call([refbin.REF_BIN, "compress", os.path.join(GOLD, "config1", db), os.path.join(out, db), "--threads", "1",
                               "-v", "1"])
        for f in os.listdir(out):       # only the DB triple is a fixture (compress also copies header / lookup side files)
            if f.startswith(db) and f not in (db, db + ".index", db + ".dbtype"):
                os.remove(os.path.join(out, f))
    print("compressed fixtures written")


if __name__ == "__main__":
    main()
