"""Synthetic spacer / phage-ORF sets of the shapes BASELINE.json names.

Targets ("GenBank_phage_2018_09-shaped"): G genomes, ORFs per genome ~ Poisson(62), ORF length
30 + Geometric(mean 77) capped at 32734, residues i.i.d. from the VTML40 background, no X.
Queries: N spacers of 25..45 nt, 30 % planted (a window of a target ORF back-translated with a fixed
codon table, 0..4 nt substitutions), six-frame translated, split at stops, fragments >= 9 aa kept
(what createsetdb --extractorf-spacer 1 + translatenucs produce, data/createsetdb.sh:88-113).

All generators are deterministic functions of their seed and are vectorised with numpy so that the
1 M spacer configuration is generated in seconds.
"""
import numpy as np

AA = "ACDEFGHIKLMNPQRSTVWYX"

# VTML40 background frequencies (lib/mmseqs/data/VTML40.out header), 20 residues
BG40 = np.array([0.0819, 0.0169, 0.0453, 0.0616, 0.0465, 0.0745, 0.0217, 0.0549, 0.0582, 0.0905,
                 0.0274, 0.0435, 0.0506, 0.0375, 0.0447, 0.0533, 0.0569, 0.081, 0.0132, 0.0329])

# standard genetic code, codon index = 16*b0 + 4*b1 + b2 with A,C,G,T = 0..3
_CODE = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF"
_AA_OF_CODON = np.array([AA.index(c) if c != "*" else 255 for c in _CODE], dtype=np.uint8)
# fixed back-translation table: first codon (in index order) of each amino acid
_CODON_OF_AA = np.zeros(21, dtype=np.int64)
for _a in range(20):
    _CODON_OF_AA[_a] = int(np.nonzero(_AA_OF_CODON == _a)[0][0])
_CODON_OF_AA[20] = 0


def make_targets(n_genomes=8000, seed=20180901, orfs_per_genome=62.0, mean_extra=77.0, max_len=32734):
    """Returns (residues uint8, offsets uint64[T+1], set_of_orf uint32[T])."""
    rng = np.random.default_rng(seed)
    n_orf = rng.poisson(orfs_per_genome, size=n_genomes)
    n_orf = np.maximum(n_orf, 1)
    T = int(n_orf.sum())
    lens = 30 + rng.geometric(1.0 / mean_extra, size=T)
    lens = np.minimum(lens, max_len).astype(np.uint64)
    off = np.zeros(T + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    p = BG40 / BG40.sum()
    res = rng.choice(20, size=int(off[-1]), p=p).astype(np.uint8)
    set_of = np.repeat(np.arange(n_genomes, dtype=np.uint32), n_orf)
    return res, off, set_of


def reverse_targets(res, off):
    """Control set: every ORF reversed at the protein level (src/util/ReverseSeqByCodon.cpp:40-53)."""
    out = np.empty_like(res)
    T = len(off) - 1
    lens = np.diff(off).astype(np.int64)
    start = off[:-1].astype(np.int64)
    idx = np.arange(len(res), dtype=np.int64)
    tid = np.repeat(np.arange(T, dtype=np.int64), lens)
    local = idx - start[tid]
    out[idx] = res[start[tid] + (lens[tid] - 1 - local)]
    return out


def make_spacers(n_spacers, tres, toff, seed=None, planted=0.3, min_nt=25, max_nt=45, max_subst=4, genome_of=None, set_of=None):
    """Nucleotide spacers as a padded array: returns (nt uint8[n, max_nt] with 255 padding, lens).
    genome_of (int64[n], -1 = none) + set_of (genome of every ORF, ascending): planted spacers copy an ORF of THAT genome
    (spacers of one CRISPR array hitting one phage); spacers with genome -1 stay random."""
    if seed is None:
        seed = 42 + n_spacers
    rng = np.random.default_rng(seed)
    lens = rng.integers(min_nt, max_nt + 1, size=n_spacers)
    nt = rng.integers(0, 4, size=(n_spacers, max_nt)).astype(np.uint8)
    is_planted = rng.random(n_spacers) < planted
    T = len(toff) - 1
    pl = np.nonzero(is_planted)[0]
    if len(pl) and T > 0:
        tlen = np.diff(toff).astype(np.int64)
        naa = (max_nt + 2) // 3 + 1  # residues needed to cover the window in any codon phase
        ok = np.nonzero(tlen >= naa)[0]
        tsel = ok[rng.integers(0, len(ok), size=len(pl))]
        if genome_of is not None:
            g = np.asarray(genome_of, dtype=np.int64)[pl]
            first = np.searchsorted(set_of, np.maximum(g, 0), side="left")
            cnt = np.searchsorted(set_of, np.maximum(g, 0), side="right") - first
            pick = first + (rng.random(len(pl)) * np.maximum(cnt, 1)).astype(np.int64)
            usable = (g >= 0) & (cnt > 0) & (tlen[np.minimum(pick, T - 1)] >= naa)
            tsel = np.where(usable, np.minimum(pick, T - 1), tsel)
            keep = usable
            pl, tsel = pl[keep], tsel[keep]
        start_aa = (rng.random(len(pl)) * (tlen[tsel] - naa + 1)).astype(np.int64)
        phase = rng.integers(0, 3, size=len(pl))
        aa = tres[(toff[tsel].astype(np.int64) + start_aa)[:, None] + np.arange(naa)[None, :]]
        cod = _CODON_OF_AA[aa]  # [n, naa]
        bases = np.stack([cod // 16, (cod // 4) % 4, cod % 4], axis=2).reshape(len(pl), naa * 3)
        win = bases[np.arange(len(pl))[:, None], phase[:, None] + np.arange(max_nt)[None, :]].astype(np.uint8)
        nsub = rng.integers(0, max_subst + 1, size=len(pl))
        for k in range(max_subst):
            m = nsub > k
            pos = (rng.random(len(pl)) * lens[pl]).astype(np.int64)
            newb = rng.integers(0, 4, size=len(pl)).astype(np.uint8)
            rows = np.nonzero(m)[0]
            win[rows, pos[rows]] = newb[rows]
        strand = rng.random(len(pl)) < 0.5
        # reverse-complement (A<->T, C<->G with A,C,G,T = 0..3) the first len nucleotides of half the spacers
        for L in range(min_nt, max_nt + 1):
            rows = np.nonzero(strand & (lens[pl] == L))[0]
            if len(rows):
                win[rows, :L] = (3 - win[rows, :L][:, ::-1])
        nt[pl] = win
    nt[np.arange(max_nt)[None, :] >= lens[:, None]] = 255
    return nt, lens.astype(np.int64)


def six_frame_fragments(nt, lens, min_aa=9, return_loc=False):
    """Six-frame translation, split at stop codons, keep fragments >= min_aa.

    Returns (qres uint8, qoff uint64[Q+1], spacer_of_frame uint32[Q]).  Frame order per spacer:
    forward frames 1,2,3 then reverse frames 1,2,3; fragments left to right (extractorfs order).
    return_loc: additionally int64[Q, 3] = (strand * 3 + frame, first codon, codons) of every fragment."""
    n, W = nt.shape
    comp = nt.copy()
    valid = comp != 255
    comp[valid] = 3 - comp[valid]
    # reverse complement aligned to the left
    rc = np.full_like(nt, 255)
    for L in np.unique(lens):
        rows = np.nonzero(lens == L)[0]
        rc[rows, :L] = comp[rows, :L][:, ::-1]
    out_chunks = []
    for strand, arr in ((0, nt), (1, rc)):
        for f in range(3):
            ncod = (W - f) // 3
            if ncod <= 0:
                continue
            b = arr[:, f:f + 3 * ncod].reshape(n, ncod, 3).astype(np.int64)
            full = (b != 255).all(axis=2)
            idx = np.where(full, 16 * b[:, :, 0] + 4 * b[:, :, 1] + b[:, :, 2], 0)
            aa = np.where(full, _AA_OF_CODON[idx], 254)  # 255 = stop, 254 = beyond end
            out_chunks.append((strand * 3 + f, aa))
    # split every (spacer, frame) row at stops; done with run-length logic on a flattened view
    order_res, order_off, order_sp, order_key, order_loc = [], [], [], [], []
    for fidx, aa in out_chunks:
        n_, C_ = aa.shape
        is_res = aa < 21
        # run starts / ends per row
        pad = np.zeros((n_, 1), dtype=bool)
        x = np.concatenate([pad, is_res, pad], axis=1)
        d = np.diff(x.astype(np.int8), axis=1)
        srow, scol = np.nonzero(d == 1)
        erow, ecol = np.nonzero(d == -1)
        flen = ecol - scol
        keep = flen >= min_aa
        srow, scol, flen = srow[keep], scol[keep], flen[keep]
        order_sp.append(srow)
        order_loc.append(np.stack([np.full(len(srow), fidx, dtype=np.int64), scol.astype(np.int64), flen.astype(np.int64)], axis=1))
        order_key.append(srow.astype(np.int64) * 1000 + fidx * 100 + scol)
        order_res.append((aa, srow, scol, flen))
    sp = np.concatenate(order_sp) if order_sp else np.zeros(0, dtype=np.int64)
    key = np.concatenate(order_key) if order_key else np.zeros(0, dtype=np.int64)
    flen_all = np.concatenate([r[3] for r in order_res]) if order_res else np.zeros(0, dtype=np.int64)
    perm = np.argsort(key, kind="stable")
    Q = len(perm)
    qoff = np.zeros(Q + 1, dtype=np.uint64)
    np.cumsum(flen_all[perm], out=qoff[1:])
    qres = np.empty(int(qoff[-1]), dtype=np.uint8)
    # scatter fragments
    base = 0
    inv = np.empty(Q, dtype=np.int64)
    inv[perm] = np.arange(Q)
    for aa, srow, scol, flen in order_res:
        m = len(srow)
        if m == 0:
            continue
        dst0 = qoff[inv[base:base + m]].astype(np.int64)
        maxl = int(flen.max())
        cols = np.arange(maxl)[None, :]
        mask = cols < flen[:, None]
        src = aa[srow[:, None], np.minimum(scol[:, None] + cols, aa.shape[1] - 1)]
        qres[(dst0[:, None] + cols)[mask]] = src[mask].astype(np.uint8)
        base += m
    if return_loc:
        loc = np.concatenate(order_loc) if order_loc else np.zeros((0, 3), dtype=np.int64)
        return qres, qoff, sp[perm].astype(np.uint32), loc[perm]
    return qres, qoff, sp[perm].astype(np.uint32)


_NT = np.frombuffer(b"ACGT", dtype=np.uint8)


def fragment_nucleotides(nt, lens, sp, loc):
    """ASCII nucleotides of every fragment of six_frame_fragments(..., return_loc=True): the codons the fragment was
    translated from, on its strand (what <db>_nucl_orf holds next to <db>, data/createsetdb.sh:88-113)."""
    out = []
    for q in range(len(sp)):
        row, (fidx, scol, flen) = int(sp[q]), (int(x) for x in loc[q])
        L = int(lens[row])
        a = nt[row, :L]
        if fidx >= 3:
            a = (3 - a)[::-1]
        f = fidx % 3
        out.append(_NT[a[f + 3 * scol:f + 3 * (scol + flen)]].tobytes())
    return out


def back_translate(res, off, reverse_codons=False):
    """ASCII coding sequence of every ORF under the fixed codon table (_CODON_OF_AA); reverse_codons: codon order
    reversed, every codon intact (reverseseqbycodon, src/util/ReverseSeqByCodon.cpp:40-53 -- translates to the reversed ORF)."""
    cod = _CODON_OF_AA[np.asarray(res)]
    bases = _NT[np.stack([cod // 16, (cod // 4) % 4, cod % 4], axis=1)]
    out = []
    for i in range(len(off) - 1):
        b = bases[int(off[i]):int(off[i + 1])]
        out.append((b[::-1] if reverse_codons else b).tobytes())
    return out


def make_spacer_arrays(n_arrays, tres, toff, set_of, seed=7, min_spacers=2, max_spacers=25, hit_arrays=0.4, planted=0.5,
                       return_nucl=False, **kw):
    """CRISPR arrays (query SETS): array a holds U{min..max} spacers; a fraction hit_arrays of the arrays targets one random
    genome (their planted spacers copy ORFs of it).  Returns (qres, qoff, set_of_frame uint32[Q], n_arrays): the frames of
    all spacers (six-frame, as make_query_frames) with the array every frame belongs to."""
    rng = np.random.default_rng(seed)
    n_sp = rng.integers(min_spacers, max_spacers + 1, size=n_arrays)
    array_of = np.repeat(np.arange(n_arrays, dtype=np.int64), n_sp)
    G = int(set_of[-1]) + 1 if len(set_of) else 0
    genome = np.where(rng.random(n_arrays) < hit_arrays, rng.integers(0, max(G, 1), size=n_arrays), -1)
    nt, lens = make_spacers(int(n_sp.sum()), tres, toff, seed=seed + 1, planted=planted, genome_of=genome[array_of], set_of=set_of, **kw)
    if return_nucl:
        qres, qoff, sp, loc = six_frame_fragments(nt, lens, return_loc=True)
        return qres, qoff, array_of[sp].astype(np.uint32), n_arrays, fragment_nucleotides(nt, lens, sp, loc)
    qres, qoff, sp = six_frame_fragments(nt, lens)
    return qres, qoff, array_of[sp].astype(np.uint32), n_arrays


def write_array_search_dbs(work, n_genomes=150, target_seed=20180901, n_arrays=140, array_seed=11, nucl=False):
    """BASELINE config 3 at a chosen scale as MMseqs2 set DBs under `work`: `targets` (phage ORF setDB), `targets_rev`
    (reversed-ORF control setDB) and `qsetdb` (frames of synthetic CRISPR arrays), each with its _member_to_set /
    _set_to_member / _set_size files.  Returns the SHA-256 over the generated arrays (pinned next to the golden outputs)."""
    import hashlib
    import os
    from . import dbio
    tres, toff, set_of = make_targets(n_genomes=n_genomes, seed=target_seed)
    rres = reverse_targets(tres, toff)
    qres, qoff, qset, _, qnt = make_spacer_arrays(n_arrays, tres, toff, set_of, seed=array_seed, return_nucl=True)
    if nucl:
        # the nucleotide ORF DBs of the --perform-nucl-aln 1 branch (same keys as the amino-acid DBs)
        for name, seqs in (("targets", back_translate(tres, toff)), ("targets_rev", back_translate(rres, toff)), ("qsetdb", qnt)):
            dbio.write_db(os.path.join(work, name + "_nucl_orf"), ((i, s + b"\n") for i, s in enumerate(seqs)), dbio.DBTYPE_NUCLEOTIDES)
    dbio.write_set_db(os.path.join(work, "targets"), tres, toff, set_of)
    dbio.write_set_db(os.path.join(work, "targets_rev"), rres, toff, set_of)
    dbio.write_set_db(os.path.join(work, "qsetdb"), qres, qoff, qset)
    h = hashlib.sha256()
    for a in (tres, toff, set_of, rres, qres, qoff, qset):
        h.update(a.tobytes())
    return h.hexdigest()


def make_query_frames(n_spacers, tres, toff, seed=None, **kw):
    nt, lens = make_spacers(n_spacers, tres, toff, seed=seed, **kw)
    return six_frame_fragments(nt, lens)


def decode(res):
    return "".join(AA[int(a)] for a in res)


def write_fasta(path, res, off, prefix="s"):
    lut = np.frombuffer(AA.encode(), dtype=np.uint8)
    txt = lut[res]
    with open(path, "wb") as f:
        for i in range(len(off) - 1):
            f.write(b">%s%d\n" % (prefix.encode(), i))
            f.write(txt[int(off[i]):int(off[i + 1])].tobytes())
            f.write(b"\n")
