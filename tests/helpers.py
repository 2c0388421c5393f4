import zlib

import numpy as np


def crc(x):
    return "%08x" % zlib.crc32(np.ascontiguousarray(x).astype("<i4").tobytes())


def gen_case(bs, case):
    return getattr(bs, case["gen"])(*case["args"])


def random_csr(rng, n_rows, n_cols, avg, sort=True, dups=False):
    """Random boolean CSR; optionally unsorted rows with repeated columns (legal input for the reference)."""
    lens = rng.poisson(avg, n_rows).astype(np.int64)
    row = np.zeros(n_rows + 1, np.int32)
    cols = []
    for i in range(n_rows):
        c = rng.integers(0, n_cols, lens[i]) if n_cols > 0 else np.zeros(0, np.int64)
        if not dups:
            c = np.unique(c)
        if sort:
            c = np.sort(c)
        cols.append(c)
        row[i + 1] = row[i] + len(c)
    col = np.concatenate(cols).astype(np.int32) if cols else np.zeros(0, np.int32)
    return row, col


def assert_csr_contract(Ccol, Crow, Bm):
    """Result contract: Crow[0]=0, monotone, columns strictly ascending per row, in [0,Bm)."""
    Crow = np.asarray(Crow, dtype=np.int64)
    assert Crow[0] == 0
    assert (np.diff(Crow) >= 0).all()
    assert Crow[-1] == len(Ccol)
    if len(Ccol):
        assert Ccol.min() >= 0 and Ccol.max() < Bm
        d = np.diff(Ccol.astype(np.int64))
        starts = Crow[1:-1]
        starts = starts[(starts > 0) & (starts < len(Ccol))]
        ok = d > 0
        ok[starts - 1] = True            # a row boundary may go down
        assert ok.all()
