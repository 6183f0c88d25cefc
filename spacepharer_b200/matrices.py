"""Substitution matrices: embedded tables rendered to the `.out` text layout the parsers accept.

The reference embeds its matrix files in the binary and falls back to them when `--sub-mat NAME.out`
is not a readable path (M/commons/Parameters.cpp:14-18, SubstitutionMatrix.cpp:13-36); this module
and csrc/host/submat.cpp do the same with the tables in matrices_data.py / matrix_data.inc.
"""
import os

from .matrices_data import MATRICES


def matrix_text(name):
    """`name` = embedded matrix name (e.g. 'VTML40.out') or a path to a .out file."""
    if os.path.isfile(name):
        with open(name) as f:
            return f.read()
    base = os.path.basename(name)
    if base not in MATRICES:
        raise KeyError("unknown substitution matrix %r (embedded: %s)" % (name, ", ".join(sorted(MATRICES))))
    m = MATRICES[base]
    lines = ["# %s" % base,
             "# Background (precomputed optional): %s" % m["background"],
             "# Lambda     (precomputed optional): %s" % m["lambda"],
             "\t" + "\t".join(m["alphabet"])]
    for c, row in zip(m["alphabet"], m["scores"]):
        lines.append(c + "\t" + "\t".join(row.split()))
    return "\n".join(lines) + "\n"
