"""Driver for the reference binary built by `make -C oracle ref` (oracle/_ref/spacepharer).

TEST / BASELINE INFRASTRUCTURE ONLY: imported by bench.py's `--impl reference` arm and `cpu_baseline` leg and by the golden
generators under tools/; never by the product (spacepharer_b200/).  The binary is the UNMODIFIED reference
(soedinglab/spacepharer, AVX2 + OpenMP build); this module only writes MMseqs2 DBs from the synthetic workload, runs the two
modules of the hot path with the argv the reference's own workflow generates (`blastp.sh:33,41`), and parses the
times the modules print themselves.
"""
import os
import re
import shutil
import subprocess
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "spacepharer")

PREF_FLAGS = ("--sub-mat nucl:nucleotide.out,aa:VTML40.out --seed-sub-mat nucl:nucleotide.out,aa:VTML80.out -k 6 "
              "--k-score 2147483647 --alph-size nucl:5,aa:21 --max-seq-len 65535 --max-seqs {maxseqs} --split 0 "
              "--split-mode 2 --split-memory-limit 0 -c 0 --cov-mode 0 --comp-bias-corr 1 --diag-score 1 "
              "--exact-kmer-matching 0 --mask {mask} --mask-lower-case 0 --min-ungapped-score 15 --add-self-matches 0 "
              "--spaced-kmer-mode 1 --db-load-mode 0 --pca 1 --pcb 1.5 --spaced-kmer-pattern 11011101 --threads {threads} "
              "--compressed 0 -v 3 -s {sens}")
ALN_FLAGS = ("--sub-mat nucl:nucleotide.out,aa:VTML40.out -a 1 --alignment-mode 2 --alignment-output-mode 0 "
             "--wrapped-scoring 0 -e {evalue} --min-seq-id 0 --min-aln-len 0 --seq-id-mode 0 --alt-ali 0 -c 0 --cov-mode 0 "
             "--max-seq-len 65535 --comp-bias-corr 1 --max-rejected 2147483647 --max-accept 2147483647 "
             "--add-self-matches 0 --db-load-mode 0 --pca 1 --pcb 1.5 --score-bias 0 --realign 0 "
             "--realign-score-bias -0.2 --realign-max-seqs 2147483647 --gap-open nucl:10,aa:{go} --gap-extend {ge} "
             "--zdrop 40 --threads {threads} --compressed 0 -v 3")

_TIME = r"(\d+)h (\d+)m (\d+)s (\d+)ms"


def available():
    return os.path.isfile(REF_BIN) and os.access(REF_BIN, os.X_OK)


def _module_times(text):
    """'Time for processing' and 'Time for index table init' as the module prints them (Timer::lap, M/commons/Timer.h)."""
    proc = re.findall(r"Time for processing: " + _TIME, text)
    init = re.findall(r"Time for index table init: " + _TIME, text)
    to_s = lambda g: int(g[0]) * 3600.0 + int(g[1]) * 60.0 + int(g[2]) + int(g[3]) / 1000.0
    if not proc:
        raise RuntimeError("the reference module printed no 'Time for processing' line:\n" + text[-2000:])
    return to_s(proc[-1]), (to_s(init[-1]) if init else 0.0)


def write_sequence_db(name, res, off):
    """numeric residues (0..20) + offsets -> amino-acid DB `name` with keys 0..n-1 in order (createdb --shuffle 0 layout)."""
    from spacepharer_b200 import dbio
    lut = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYX", dtype=np.uint8)
    txt = lut[np.asarray(res)]
    n = len(off) - 1
    dbio.write_sequence_db(name, range(n), (txt[int(off[i]):int(off[i + 1])].tobytes() for i in range(n)))


class Workdir:
    """Temporary directory holding the target DB and query DB(s) of one workload."""

    def __init__(self, tres, toff, prefix="spb_ref_"):
        self.path = tempfile.mkdtemp(prefix=prefix, dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        write_sequence_db(os.path.join(self.path, "targets"), tres, toff)
        self.n = 0

    def add_queries(self, qres, qoff, name="queries"):
        write_sequence_db(os.path.join(self.path, name), qres, qoff)
        return name

    def close(self):
        shutil.rmtree(self.path, ignore_errors=True)


def run_module(args, cwd):
    r = subprocess.run([REF_BIN] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference %s failed (rc %d):\n%s" % (args[0], r.returncode, r.stdout[-2000:]))
    return r.stdout


def search(work, queries="queries", threads=None, tag="r", max_seqs=300, mask=1, sens=5.7, evalue=200, go=16, ge=2, keep=False):
    """One pass of the hot path with the reference: `prefilter` then `align` on the DBs in `work`.
    -> dict(prefilter_s = module 'Time for processing' minus 'Time for index table init' (the k-mer matching itself; the GPU arm
    keeps its index resident too), index_s, align_s, wall_s, step_s = prefilter_s + align_s)."""
    threads = threads or (os.cpu_count() or 1)
    w = work.path if isinstance(work, Workdir) else work
    pref, aln = "pref_" + tag, "aln_" + tag
    for stem in (pref, aln):
        for f in os.listdir(w):
            if f == stem or f.startswith(stem + "."):
                os.remove(os.path.join(w, f))
    t0 = time.perf_counter()
    out = run_module(["prefilter", queries, "targets", pref] +
                     PREF_FLAGS.format(maxseqs=max_seqs, mask=mask, threads=threads, sens=sens).split(), w)
    p_total, p_init = _module_times(out)
    out = run_module(["align", queries, "targets", pref, aln] +
                     ALN_FLAGS.format(evalue=evalue, go=go, ge=ge, threads=threads).split(), w)
    a_total, _ = _module_times(out)
    wall = time.perf_counter() - t0
    m = re.search(r"(\d+) sequence pairs passed the thresholds", out)
    res = {"prefilter_s": max(p_total - p_init, 1e-6), "index_s": p_init, "align_s": a_total, "wall_s": wall,
           "step_s": max(p_total - p_init, 1e-6) + a_total, "rows": int(m.group(1)) if m else -1, "threads": threads}
    if not keep:
        for stem in (pref, aln):
            for f in os.listdir(w):
                if f == stem or f.startswith(stem + "."):
                    os.remove(os.path.join(w, f))
    return res
