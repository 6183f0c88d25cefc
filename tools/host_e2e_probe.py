#!/usr/bin/env python
"""Wall-clock of the host binary on BASELINE config 4 (1 M spacers): `spb200 search` from on-disk DBs to the on-disk alignment DB
(development aid; prints the binary's own log and the total)."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gen_golden_config2 import write_config2
from test_host_cli import ALN_FLAGS, BIN, PREF_FLAGS
w = tempfile.mkdtemp()
t0 = time.time(); nq, nt = write_config2(w, int(sys.argv[1]) if len(sys.argv) > 1 else 1000000); print("inputs written: %d frames, %d ORFs, %.1f s" % (nq, nt, time.time() - t0))
threads = str(min(16, os.cpu_count() or 1))
seen, merged = set(), []
for lst in (PREF_FLAGS, ALN_FLAGS):
    for i in range(0, len(lst), 2):
        if lst[i] not in seen:
            seen.add(lst[i]); merged += [lst[i], threads if lst[i] == "--threads" else lst[i + 1]]
for extra in (["--batch", "250000", "--lanes", "1"], ["--batch", "250000", "--lanes", "2"], ["--batch", "250000", "--lanes", "1"], ["--batch", "250000", "--lanes", "2"]):
    t0 = time.time()
    r = subprocess.run([BIN, "search", w + "/queries", w + "/targets", w + "/aln"] + merged + extra, capture_output=True, text=True,
                       env=dict(os.environ, SPB_TIMING="1"))
    dt = time.time() - t0
    print("spb200 search %s: %.2f s wall (%d threads), rc %d, %.0f frames/s from disk to disk" % (" ".join(extra) or "(one 1 M-frame batch per GPU pass)", dt, int(threads), r.returncode, nq / dt))
    print("  " + "\n  ".join(l for l in r.stdout.split("\n") if l.strip())[:1200])
    print("  " + "\n  ".join(l for l in r.stderr.split("\n") if "timing" in l))
