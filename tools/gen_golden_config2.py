#!/usr/bin/env python
"""Full-size golden digest: the REFERENCE binary's prefilter and alignment DBs for BASELINE config 2 (10 000 synthetic spacers =
38 024 frames vs the 495 567-ORF synthetic GenBank-phage-shaped set; spacepharer_b200.synth, fixed seeds), reduced to SHA-256
digests of their canonical form (entries in key order) so that the fixture stays a few hundred bytes.  tests/test_gpu_fullsize.py
regenerates the inputs, runs spb200 prefilter + align with the same argv and compares the digests: byte identity with the
reference at the size bench.py measures.  Development time only (needs the reference build, tools/gen_golden.py).
Usage: python tools/gen_golden_config2.py [--ref-bin /tmp/refbuild/src/spacepharer]"""
import argparse
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gen_golden import ALN_FLAGS, PREF_FLAGS, run  # noqa: E402
from spacepharer_b200 import dbio, synth  # noqa: E402


def write_config2(work, spacers=10000):
    tres, toff, _ = synth.make_targets(n_genomes=8000)
    qres, qoff, _ = synth.make_query_frames(spacers, tres, toff)
    lut = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYX", dtype=np.uint8)
    for name, res, off in (("targets", tres, toff), ("queries", qres, qoff)):
        txt = lut[res]
        n = len(off) - 1
        dbio.write_sequence_db(os.path.join(work, name), range(n), (txt[int(off[i]):int(off[i + 1])].tobytes() for i in range(n)))
    return len(qoff) - 1, len(toff) - 1


def db_digest(name):
    """SHA-256 over `key\\n` + entry bytes for every key in ascending order; also the number of non-empty lines"""
    h = hashlib.sha256()
    rows = 0
    ent = dbio.read_entries(name)
    for k in sorted(ent):
        h.update(b"%d\n" % k)
        h.update(ent[k])
        rows += ent[k].count(b"\n")
    return h.hexdigest(), rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-bin", default="/tmp/refbuild/src/spacepharer")
    ap.add_argument("--spacers", type=int, default=10000, help="10000 = config 2, 1000000 = config 4 (writes config4_digest.json)")
    ap.add_argument("--out", default=None)
    ap.add_argument("--threads", type=int, default=4)
    a = ap.parse_args()
    if a.out is None:
        a.out = os.path.join(ROOT, "tests", "golden", "config2_digest.json" if a.spacers == 10000 else "config%s_digest.json" %
                             ("4" if a.spacers == 1000000 else "_%d" % a.spacers))
    w = tempfile.mkdtemp(prefix="spb_cfg2_")
    nq, nt = write_config2(w, a.spacers)
    run("%s prefilter %s/queries %s/targets %s/pref %s" % (a.ref_bin, w, w, w, PREF_FLAGS.format(maxseqs=300, mask=1).replace("--threads 4", "--threads %d" % a.threads)))
    run("%s align %s/queries %s/targets %s/pref %s/aln %s" % (a.ref_bin, w, w, w, w, ALN_FLAGS.format(evalue=200, go=16, ge=2).replace("--threads 4", "--threads %d" % a.threads)))
    pd, pr = db_digest(w + "/pref")
    ad, ar = db_digest(w + "/aln")
    import shutil
    json.dump({"spacers": a.spacers, "n_queries": nq, "n_targets": nt, "pref_sha256": pd, "pref_rows": pr, "aln_sha256": ad, "aln_rows": ar,
               "note": "reference soedinglab/spacepharer @ d1e94c0, AVX2 build, --threads 4; canonical form: entries in key order"},
              open(a.out, "w"), indent=1)
    print(open(a.out).read())
    shutil.rmtree(w)


if __name__ == "__main__":
    main()
