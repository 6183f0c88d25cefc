"""Minimal MMseqs2 on-disk DB helpers (Python side: tests, bench and the Python mirror).

Format (M/commons/DBReader.h:57-99, DBWriter.cpp:216-440): `<name>` holds NUL-terminated entries,
`<name>.index` text lines `key\\toffset\\tlength` (length includes the NUL), `<name>.dbtype` a 4-byte
little-endian int (0 amino acids, 5 alignment results, 7 prefilter results).  Multi-threaded writers
may leave `<name>.0 .. <name>.N-1` whose offsets live in the concatenated space.  DBs written with `--compressed 1`
(bit 31 of the dbtype) are inflated while reading.  The native reader /
writer used by the spb200 binary is spacepharer_b200/csrc/host/db_io.cpp.
"""
import glob
import os
import struct

import numpy as np

DBTYPE_AMINO_ACIDS = 0
DBTYPE_NUCLEOTIDES = 1
DBTYPE_ALIGNMENT_RES = 5
DBTYPE_PREFILTER_RES = 7


def read_db(name):
    """-> (keys uint32[n], offsets int64[n], lengths int64[n], data bytes); index sorted by key."""
    if os.path.exists(name):
        data = open(name, "rb").read()
    else:
        parts = sorted(glob.glob(name + ".[0-9]*"), key=lambda x: int(x.rsplit(".", 1)[1]))
        parts = [p for p in parts if p.rsplit(".", 1)[1].isdigit()]
        if not parts:
            raise FileNotFoundError(name)
        data = b"".join(open(p, "rb").read() for p in parts)
    rows = [tuple(int(x) for x in line.split()[:3]) for line in open(name + ".index") if line.strip()]
    rows.sort()
    a = np.array(rows, dtype=np.int64).reshape(-1, 3)
    if os.path.exists(name + ".dbtype") and read_dbtype(name) < 0:      # bit 31: --compressed 1
        a, data = _inflate(a, data)
    return a[:, 0].astype(np.uint32), a[:, 1].copy(), a[:, 2].copy(), data


def _inflate(rows, data):
    """Compressed DB (DBReader.cpp:512-539, DBWriter.cpp:254-397): an entry is `u32 size | size bytes | terminator`, terminator 0
    = one zstd frame, 0xFF = stored; the index length is the uncompressed length + 1.  -> rows and data in the uncompressed
    layout (entries in offset order).  zstd through the system's libzstd.so.1, like the host binary."""
    import ctypes as C
    z = C.CDLL("libzstd.so.1")
    z.ZSTD_decompress.restype = C.c_size_t
    z.ZSTD_decompress.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    z.ZSTD_isError.argtypes = [C.c_size_t]
    out, pos = [], 0
    rows = rows.copy()
    for i in np.argsort(rows[:, 1], kind="stable"):
        o, want = int(rows[i, 1]), int(rows[i, 2]) - 1
        size = struct.unpack_from("<I", data, o)[0]
        body = data[o + 4:o + 4 + size]
        if data[o + 4 + size] == 0:
            buf = C.create_string_buffer(max(want, 1))
            got = z.ZSTD_decompress(buf, want, body, size)
            if z.ZSTD_isError(got) or got != want:
                raise ValueError("could not decompress the entry at offset %d" % o)
            body = buf.raw[:want]
        elif size != want:
            raise ValueError("stored entry at offset %d: %d bytes, index says %d" % (o, size, want))
        out.append(body + b"\0")
        rows[i, 1] = pos
        pos += want + 1
    return rows, b"".join(out)


def read_dbtype(name):
    with open(name + ".dbtype", "rb") as f:
        return struct.unpack("<i", f.read(4))[0]


def read_entries(name):
    keys, off, ln, data = read_db(name)
    return {int(k): data[o:o + l - 1] for k, o, l in zip(keys, off, ln)}


def read_sequence_db(name):
    """Sequence DB -> (keys in LINEAR_ACCESS order (by offset, DBReader.cpp:339-368), list of ASCII residue bytes).
    Entries are `residues\\n\\0`, seqLen = length - 2 (DBReader.h:191-213)."""
    keys, off, ln, data = read_db(name)
    order = np.argsort(off, kind="stable")
    seqs = [data[int(off[i]):int(off[i]) + int(ln[i]) - 2] for i in order]
    return keys[order], seqs


def amino_acid_db_size(name):
    """DBReader::getAminoAcidDBSize (DBReader.cpp:541-550): data bytes - 2 * entries."""
    keys, off, ln, data = read_db(name)
    return int(ln.sum()) - 2 * len(keys)


def write_db(name, entries, dbtype):
    """entries: iterable of (key, payload bytes without NUL) in write order."""
    pos = 0
    idx = []
    with open(name, "wb") as f:
        for k, payload in entries:
            f.write(payload)
            f.write(b"\0")
            idx.append((int(k), pos, len(payload) + 1))
            pos += len(payload) + 1
    idx.sort()
    with open(name + ".index", "w") as f:
        for k, o, l in idx:
            f.write("%d\t%d\t%d\n" % (k, o, l))
    with open(name + ".dbtype", "wb") as f:
        f.write(struct.pack("<i", dbtype))


def write_sequence_db(name, keys, seqs):
    write_db(name, ((k, s + b"\n") for k, s in zip(keys, seqs)), DBTYPE_AMINO_ACIDS)


DBTYPE_GENERIC_DB = 12


def write_sequence_db_arrays(name, res, off):
    """Sequence DB (keys 0..n-1, entries `residues\\n\\0`) straight from the encoded arrays (residues uint8, offsets [n+1]):
    vectorised, for the million-entry DBs of the benchmarks (write_sequence_db takes 10 s per million entries)."""
    lut = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYX", dtype=np.uint8)
    off = np.asarray(off, dtype=np.int64)
    n = len(off) - 1
    lens = np.diff(off)
    out = np.empty(int(off[-1]) + 2 * n, dtype=np.uint8)
    seq_of = np.repeat(np.arange(n, dtype=np.int64), lens)
    out[np.arange(int(off[-1]), dtype=np.int64) + 2 * seq_of] = lut[np.asarray(res)[: int(off[-1])]]
    ends = off[1:] + 2 * np.arange(n, dtype=np.int64)
    out[ends] = 10
    out[ends + 1] = 0
    with open(name, "wb") as f:
        f.write(out.tobytes())
    starts = off[:-1] + 2 * np.arange(n, dtype=np.int64)
    import pandas as pd
    pd.DataFrame({"k": np.arange(n, dtype=np.int64), "o": starts, "l": lens + 2}).to_csv(name + ".index", sep="\t", header=False, index=False)
    with open(name + ".dbtype", "wb") as f:
        f.write(struct.pack("<i", DBTYPE_AMINO_ACIDS))


def write_set_db(name, res, off, set_of):
    """A set database like `createsetdb` leaves it for the search and the set-level stages (data/createsetdb.sh:115-160):
    the amino-acid sequence DB <name> (key = sequence index) plus <name>_member_to_set (sequence -> set key),
    <name>_set_to_member (set -> its sequences, one per line) and <name>_set_size (set -> number of sequences)."""
    import numpy as np
    lut = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYX", dtype=np.uint8)
    txt = lut[np.asarray(res)]
    n = len(off) - 1
    write_sequence_db(name, range(n), (txt[int(off[i]):int(off[i + 1])].tobytes() for i in range(n)))
    set_of = np.asarray(set_of, dtype=np.int64)
    write_db(name + "_member_to_set", ((i, b"%d\n" % set_of[i]) for i in range(n)), DBTYPE_ALIGNMENT_RES)
    order = np.argsort(set_of, kind="stable")
    sets, first = np.unique(set_of[order], return_index=True)
    bounds = list(first) + [n]
    write_db(name + "_set_to_member",
             ((int(s), b"".join(b"%d\n" % m for m in order[bounds[j]:bounds[j + 1]])) for j, s in enumerate(sets)), DBTYPE_ALIGNMENT_RES)
    write_db(name + "_set_size", ((int(s), b"%d\n" % (bounds[j + 1] - bounds[j])) for j, s in enumerate(sets)), DBTYPE_GENERIC_DB)
