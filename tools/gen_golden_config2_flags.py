#!/usr/bin/env python
"""Full-size digests for the flag surface beyond the predictmatch defaults: the REFERENCE binary (oracle/_ref/spacepharer) on the
BASELINE config 2 DBs (tools/gen_golden_config2.py: 38 024 frames vs 495 567 ORFs) with
  * align   -c 0.8 --cov-mode 2 --min-seq-id 0.7 --min-aln-len 8          (acceptance criteria)
  * align   -a 0 --alignment-mode 3 -c 0.8 --cov-mode 2                    (rows without backtrace)
  * align   --seq-id-mode 1 --min-seq-id 0.9
  * prefilter --exact-kmer-matching 1
  * prefilter --split 2 --split-mode 0                                      (the reference's own lossy target split)
reduced to SHA-256 digests (tests/golden/config2_flags_digest.json).  tests/test_gpu_fullsize.py reproduces them with spb200.
Development time only.  Usage: python tools/gen_golden_config2_flags.py"""
import json
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gen_golden import run  # noqa: E402
from gen_golden_config2 import db_digest, write_config2  # noqa: E402
from oracle import refbin  # noqa: E402

# name -> (module, flag overrides); the base argv is the workflow's (refbin.PREF_FLAGS / ALN_FLAGS)
CASES = {
    "aln_criteria": ("align", {"-c": "0.8", "--cov-mode": "2", "--min-seq-id": "0.7", "--min-aln-len": "8"}),
    "aln_noback": ("align", {"-a": "0", "--alignment-mode": "3", "-c": "0.8", "--cov-mode": "2"}),
    "aln_idmode1": ("align", {"--seq-id-mode": "1", "--min-seq-id": "0.9"}),
    "pref_exact": ("prefilter", {"--exact-kmer-matching": "1"}),
    "pref_split2": ("prefilter", {"--split": "2", "--split-mode": "0"}),
}


def with_flags(base, over):
    out = base.split()
    for k, v in over.items():
        out[out.index(k) + 1] = v
    return " ".join(out)


def main():
    assert refbin.available(), "build the reference first: make -C oracle ref"
    B = refbin.REF_BIN
    w = tempfile.mkdtemp(prefix="spb_cfg2f_")
    nq, nt = write_config2(w)
    threads = os.cpu_count() or 4
    pref = refbin.PREF_FLAGS.format(maxseqs=300, mask=1, threads=threads, sens="5.7")
    aln = refbin.ALN_FLAGS.format(evalue=200, go=16, ge=2, threads=threads)
    run("%s prefilter %s/queries %s/targets %s/pref %s" % (B, w, w, w, pref))
    out = {"n_queries": nq, "n_targets": nt, "cases": {}}
    for name, (module, over) in CASES.items():
        if module == "align":
            run("%s align %s/queries %s/targets %s/pref %s/%s %s" % (B, w, w, w, w, name, with_flags(aln, over)))
        else:
            run("%s prefilter %s/queries %s/targets %s/%s %s" % (B, w, w, w, name, with_flags(pref, over)))
        d, rows = db_digest(os.path.join(w, name))
        out["cases"][name] = {"module": module, "flags": over, "sha256": d, "rows": rows}
    out["note"] = "reference soedinglab/spacepharer @ d1e94c0, AVX2 build; canonical form: entries in key order (gen_golden_config2.db_digest)"
    path = os.path.join(ROOT, "tests", "golden", "config2_flags_digest.json")
    json.dump(out, open(path, "w"), indent=1)
    print(open(path).read())
    shutil.rmtree(w)


if __name__ == "__main__":
    main()
