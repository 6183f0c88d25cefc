#!/usr/bin/env python
"""Query preparation at scale against the reference: 200 000 synthetic nucleotide spacers (some with ambiguous bases) through the
REFERENCE `extractorfs --min-length 9 --orf-start-mode 1` + `translatenucs`; SHA-256 digests of the nucleotide ORF DB, its header
DB and the amino-acid DB (canonical form) go to tests/golden/frames_digest.json.  tests/test_frames.py regenerates the input DB
and matches `spb200 extractframes`.  Usage: python tools/gen_golden_frames_scale.py [--ref-bin ...]"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gen_golden_config2 import db_digest  # noqa: E402
from spacepharer_b200 import dbio, synth  # noqa: E402

SHAPE = dict(n_spacers=200000, target_genomes=200, target_seed=3, spacer_seed=5, n_every=41)


def write_spacer_db(name, n_spacers, target_genomes, target_seed, spacer_seed, n_every):
    """nucleotide sequence DB (keys 0..n-1) of synthetic spacers; every n_every-th one carries an N"""
    tres, toff, _ = synth.make_targets(n_genomes=target_genomes, seed=target_seed)
    nt, lens = synth.make_spacers(n_spacers, tres, toff, seed=spacer_seed)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    rng = np.random.default_rng(spacer_seed + 1)
    pos = rng.integers(0, 25, size=n_spacers)

    def entries():
        for i in range(n_spacers):
            s = bytearray(lut[nt[i, :lens[i]]].tobytes())
            if i % n_every == 7:
                s[int(pos[i])] = ord("N")
            yield i, bytes(s) + b"\n"
    dbio.write_db(name, entries(), dbio.DBTYPE_NUCLEOTIDES)
    dbio.write_db(name + "_h", ((i, b"sp%d\n" % i) for i in range(n_spacers)), dbio.DBTYPE_GENERIC_DB)   # the reference opens it
    return n_spacers


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-bin", default="/tmp/refbuild/src/spacepharer")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "frames_digest.json"))
    a = ap.parse_args()
    w = tempfile.mkdtemp(prefix="spb_frames_scale_")
    db = os.path.join(w, "sp_nucl")
    write_spacer_db(db, **SHAPE)
    R = a.ref_bin
    for cmd in ([R, "extractorfs", db, db + "_orf", "--min-length", "9", "--orf-start-mode", "1", "--create-lookup", "0", "--threads", "8", "-v", "1"],
                [R, "translatenucs", db + "_orf", db + "_aa", "--threads", "8", "-v", "1"]):
        print("+", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    out = {"shape": SHAPE, "note": "reference soedinglab/spacepharer @ d1e94c0: extractorfs --min-length 9 --orf-start-mode 1 + translatenucs"}
    for name, path in (("nucl_orf", db + "_orf"), ("nucl_orf_h", db + "_orf_h"), ("aa", db + "_aa")):
        d, r = db_digest(path)
        out[name] = {"sha256": d, "rows": r}
    json.dump(out, open(a.out, "w"), indent=1)
    print(open(a.out).read())
    shutil.rmtree(w)


if __name__ == "__main__":
    main()
