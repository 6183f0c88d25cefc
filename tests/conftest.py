import gzip
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def read_fasta_gz(path):
    seqs = []
    with gzip.open(path, "rt") as f:
        for line in f:
            if not line.startswith(">"):
                seqs.append(line.strip())
    return seqs


def read_dump(path):
    """key\\tline dumps written by tools/gen_golden.py -> {key: [line, ...]}"""
    out = {}
    with gzip.open(path, "rt") as f:
        for line in f:
            k, rest = line.rstrip("\n").split("\t", 1)
            out.setdefault(int(k), []).append(rest)
    return out


@pytest.fixture(scope="session")
def golden():
    return GOLDEN
