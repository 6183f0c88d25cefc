#!/usr/bin/env python
"""Full predictmatch golden digest (BASELINE config 3 at a larger scale than tests/golden/sets): synthetic CRISPR arrays against a
synthetic phage setDB and its reversed-ORF control, the reference's DEFAULT pipeline (search, proteinaln2nucl,
combineprotnuclaln, besthitperset, mergeresultsbyset, combinescore for both DBs, filtermatchbyfdr; data/predictmatch.sh:27-116),
module by module with the REFERENCE binary; only SHA-256 digests of the result DBs (canonical form, entries in key order) are
stored.  tests/test_nuclaln.py regenerates the inputs and matches `spb200 predictmatch` against them.
Usage: python tools/gen_golden_config3.py [--ref-bin /tmp/refbuild/src/spacepharer]"""
import argparse
import json
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gen_golden import ALN_FLAGS, PREF_FLAGS, run  # noqa: E402
from gen_golden_config2 import db_digest  # noqa: E402
from gen_golden_nucl import nucl_branch  # noqa: E402
from gen_golden_sets import set_stages  # noqa: E402
from spacepharer_b200 import synth  # noqa: E402

SHAPE = dict(n_genomes=2000, target_seed=20180901, n_arrays=3000, array_seed=23)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-bin", default="/tmp/refbuild/src/spacepharer")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "config3_digest.json"))
    a = ap.parse_args()
    B = a.ref_bin
    w = tempfile.mkdtemp(prefix="spb_cfg3_")
    inputs = synth.write_array_search_dbs(w, nucl=True, **SHAPE)
    for t, suffix in ((w + "/targets", ""), (w + "/targets_rev", "_rev")):
        run("%s prefilter %s/qsetdb %s %s/pref%s %s" % (B, w, t, w, suffix, PREF_FLAGS.format(maxseqs=300, mask=1)))
        run("%s align %s/qsetdb %s %s/pref%s %s/prot_result%s %s" % (B, w, t, w, suffix, w, suffix, ALN_FLAGS.format(evalue=200, go=16, ge=2)))
        nucl_branch(B, w + "/qsetdb", t, "%s/prot_result%s" % (w, suffix), w, suffix)
        os.makedirs(w + "/st" + suffix)
        set_stages(B, w + "/qsetdb", t, "%s/nucl_combined%s" % (w, suffix), w + "/st" + suffix, "")
    run("%s filtermatchbyfdr %s/st/cScore %s/st_rev/cScore %s/match --fdr 0.05 --report-fdr 0 --compressed 0 --threads 4 -v 3" % (B, w, w, w))
    out = {"shape": SHAPE, "inputs_sha256": inputs,
           "note": "reference soedinglab/spacepharer @ d1e94c0; digest = SHA-256 over key + entry bytes in key order; rows = lines"}
    for name, path in (("prot_result", "prot_result"), ("nucl_result", "nucl_result"), ("result", "nucl_combined"),
                       ("result_rev", "nucl_combined_rev"), ("aggregate", "st/aggregate"), ("cScore", "st/cScore"),
                       ("cScore_rev", "st_rev/cScore"), ("match", "match")):
        d, r = db_digest(os.path.join(w, path))
        out[name] = {"sha256": d, "rows": r}
    json.dump(out, open(a.out, "w"), indent=1)
    print(open(a.out).read())
    shutil.rmtree(w)


if __name__ == "__main__":
    main()
