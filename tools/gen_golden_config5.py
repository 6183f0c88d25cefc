#!/usr/bin/env python
"""Exhaustive-search golden digest (BASELINE config 5 at medium scale): every spacer frame against EVERY ORF, no k-mer stage.
Reference side = `align` over a prefilter DB that lists all targets for every query (align ignores score and
diagonal for amino-acid searches, so this is an exact all-vs-all Smith-Waterman oracle).  Only the SHA-256 of the reference's
alignment DB is stored; tests/test_gpu_exhaustive.py matches `spb200 search --exhaustive-search 1` against it.
Usage: python tools/gen_golden_config5.py [--ref-bin /tmp/refbuild/src/spacepharer]"""
import argparse
import json
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gen_golden import ALN_FLAGS, run  # noqa: E402
from gen_golden_config2 import db_digest  # noqa: E402
from spacepharer_b200 import dbio, synth  # noqa: E402

SHAPE = dict(n_genomes=1000, target_seed=20180901, spacers=150, spacer_seed=77)


def write_config5(work, n_genomes, target_seed, spacers, spacer_seed):
    tres, toff, _ = synth.make_targets(n_genomes=n_genomes, seed=target_seed)
    qres, qoff, _ = synth.make_query_frames(spacers, tres, toff, seed=spacer_seed)
    lut = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYX", dtype=np.uint8)
    for name, res, off in (("targets", tres, toff), ("queries", qres, qoff)):
        txt = lut[res]
        n = len(off) - 1
        dbio.write_sequence_db(os.path.join(work, name), range(n), (txt[int(off[i]):int(off[i + 1])].tobytes() for i in range(n)))
    return len(qoff) - 1, len(toff) - 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-bin", default="/tmp/refbuild/src/spacepharer")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "config5_digest.json"))
    a = ap.parse_args()
    w = tempfile.mkdtemp(prefix="spb_cfg5_")
    nq, nt = write_config5(w, **SHAPE)
    allrows = b"".join(b"%d\t0\t0\n" % t for t in range(nt))
    dbio.write_db(w + "/pref", ((q, allrows) for q in range(nq)), dbio.DBTYPE_PREFILTER_RES)
    run("%s align %s/queries %s/targets %s/pref %s/aln %s" % (a.ref_bin, w, w, w, w, ALN_FLAGS.format(evalue=200, go=16, ge=2).replace("--threads 4", "--threads 8")))
    d, r = db_digest(w + "/aln")
    json.dump({"shape": SHAPE, "n_queries": nq, "n_targets": nt, "pairs": nq * nt, "aln_sha256": d, "aln_rows": r,
               "note": "reference soedinglab/spacepharer @ d1e94c0: align over an all-pairs prefilter DB"}, open(a.out, "w"), indent=1)
    print(open(a.out).read())
    shutil.rmtree(w)


if __name__ == "__main__":
    main()
