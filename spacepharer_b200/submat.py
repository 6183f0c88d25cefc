"""Substitution matrix construction on the host (Python mirror of csrc/host/submat.cpp).

Follows SubstitutionMatrix::readProbMatrix + BaseMatrix::generateSubMatrix
(M/commons/SubstitutionMatrix.cpp:327-407, M/commons/BaseMatrix.cpp:96-156): P_ab = exp(lambda S_ab) pa pb from the
file's scores, lambda and background; background for the scores re-derived as row sums with X forced to 1e-5;
score = round_half_away(bitFactor * log2(P_ab / (pa pb)) + bias).  Scalar libm calls keep the doubles identical to
the C++ side.
"""
import math

import numpy as np

ANY_BACK = 1e-5


class SubMat:
    def __init__(self, text, bit_factor, score_bias):
        alph = None
        lam = bg = None
        raw = {}
        for line in text.splitlines():
            if line.startswith("# Background (precomputed optional):"):
                bg = [float(x) for x in line.split()[4:]]
            elif line.startswith("# Lambda     (precomputed optional):"):
                lam = float(line.split()[4])
            if line.startswith("#") or not line.strip():
                continue
            w = line.split()
            if len(w) <= 1:
                continue
            if alph is None:
                alph = [x[0].upper() for x in w]
                continue
            raw[w[0][0].upper()] = [float(x) for x in w[1:1 + len(alph)]]
        if alph is None or "X" not in alph:
            raise ValueError("Please add X to your substitution matrix.")
        if lam is None or bg is None:
            raise ValueError("substitution matrix without precomputed lambda/background is not supported")
        n = len(alph)
        self.alphabet = "".join(alph)
        self.alph_size = n
        x = alph.index("X")
        S = [[raw[a][j] for j in range(n)] for a in alph]
        x_positive = any(S[x][j] > 0 or S[j][x] > 0 for j in range(n))
        pb = list(bg[:n])
        if not x_positive:
            for i in range(n - 1):
                pb[i] = pb[i] * (1.0 - pb[x])
        prob = [[math.exp(lam * S[i][j]) * pb[i] * pb[j] for j in range(n)] for i in range(n)]
        rows = []
        for i in range(n):
            acc = 0.0
            for j in range(n):
                acc += prob[i][j]
            rows.append(acc)
        rows[n - 1] = ANY_BACK
        sub = np.zeros((n, n), dtype=np.int16)
        for i in range(n):
            for j in range(n):
                v = bit_factor * math.log2(prob[i][j] / (rows[i] * rows[j])) + score_bias
                sub[i, j] = int(v - 0.5) if v < 0.0 else int(v + 0.5)
        self.sub = sub
        self.pback = np.array(pb, dtype=np.float64)
        self.prob = np.array(prob, dtype=np.float64)
        self.lam = lam
        # letter mapping (SubstitutionMatrix::setupLetterMapping, SubstitutionMatrix.cpp:257-297)
        base = {c: i for i, c in enumerate(alph)}
        tab = np.full(256, base["X"], dtype=np.uint8)
        for code in range(256):
            up = chr(code).upper()
            if up in "ATGCDEFHIKLMNPQRSVWYX" and up in base:
                tab[code] = base[up]
            elif up == "J":
                tab[code] = base["L"]
            elif up == "Z":
                tab[code] = base["E"]
            elif up == "B":
                tab[code] = base["D"]
        self.aa2num = tab

    def encode(self, seq):
        if isinstance(seq, str):
            seq = seq.encode()
        return self.aa2num[np.frombuffer(seq, dtype=np.uint8)].copy()
