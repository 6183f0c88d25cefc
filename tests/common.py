"""Shared helpers for the parity tests."""
import os

import numpy as np

from oracle import pyoracle as O
from spacepharer_b200 import dbio, matrices
from conftest import GOLDEN, read_dump, read_fasta_gz

_M = {}


def mats():
    """(m80 seed matrix, m40u ungapped matrix, m40 alignment matrix) exactly as the reference builds them
    (Prefiltering.cpp:63-67, Alignment.cpp:157-162)."""
    if not _M:
        _M["m80"] = O.parse_matrix(matrices.matrix_text("VTML80.out"), 8.0, -0.2)
        _M["m40u"] = O.parse_matrix(matrices.matrix_text("VTML40.out"), 2.0, -0.2)
        _M["m40"] = O.parse_matrix(matrices.matrix_text("VTML40.out"), 2.0, 0.0)
    return _M["m80"], _M["m40u"], _M["m40"]


def load_fasta_set(name):
    """-> (T list of uint8 arrays, Q list of uint8 arrays); DB keys are 0..n-1 in file order
    (createdb --shuffle 0)."""
    T = [O.encode(s) for s in read_fasta_gz(os.path.join(GOLDEN, name, "t.fasta.gz"))]
    Q = [O.encode(s) for s in read_fasta_gz(os.path.join(GOLDEN, name, "q.fasta.gz"))]
    return T, Q


def load_config1(target="targets"):
    m80, _, _ = mats()
    tkeys, tseqs = dbio.read_sequence_db(os.path.join(GOLDEN, "config1", target))
    qkeys, qseqs = dbio.read_sequence_db(os.path.join(GOLDEN, "config1", "qsetdb"))
    T = [O.encode(s, m80) for s in tseqs]
    Q = [O.encode(s, m80) for s in qseqs]
    return tkeys, T, qkeys, Q


def mask_targets(T):
    m80, _, _ = mats()
    return [O.tantan_mask(m80, t)[0] for t in T]


def db_res_count(T):
    return int(sum(len(t) for t in T))


def pref_rows(hits, keys=None):
    """[(id, score, diag)] -> reference text rows"""
    return ["%d\t%d\t%d" % (int(keys[i]) if keys is not None else i, s, d) for i, s, d in hits]


def dump(name, fname):
    return read_dump(os.path.join(GOLDEN, name, fname))
