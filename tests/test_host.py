"""Host C surface (libbspgemm_host.so): readCOO / coo2csc / mmio-compatible reader / generators / stats.
Mirrors how the reference's drivers use them (final/SpGEMM_mpi_omp.c:307-309, :330-333).  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import crc


def _write_fixture_mtx(path, f):
    with open(path, "w") as fh:
        fh.write("%%MatrixMarket matrix coordinate pattern general\n% Generated 15-Oct-2021\n")
        fh.write(f"{int(f['M'])} {int(f['N'])} {len(f['I'])}\n")
        for i, j in zip(f["I"].tolist(), f["J"].tolist()):
            fh.write(f"{i + 1} {j + 1}\n")


def test_readCOO_fixture_golden(bs, fixture_npz, tmp_path):
    p = tmp_path / "validity_test.mtx"
    _write_fixture_mtx(p, fixture_npz)
    row, col, M, N, nnz = bs.readCOO(str(p))
    assert (M, N, nnz) == (50000, 50000, 25000)
    assert crc(row) == "0327ec62" and crc(col) == "3d773ff2"          # SURVEY.md §4 golden values
    assert (row == fixture_npz["Arow"]).all() and (col == fixture_npz["Acol"]).all()


def test_readCOO_transposes_on_read(bs, tmp_path):
    p = tmp_path / "t.mtx"
    # file matrix entries (row, col): (1,2) (3,2) (2,3) (1,1)  -> pointers by file column, indices = file rows, stable
    p.write_text("%%MatrixMarket matrix coordinate pattern general\n%c\n\n3 3 4\n1 2\n3 2\n2 3\n1 1\n")
    row, col, M, N, nnz = bs.readCOO(str(p))
    assert row.tolist() == [0, 1, 3, 4] and col.tolist() == [0, 0, 2, 1]


def test_readCOO_value_columns_are_skipped(bs, tmp_path):
    p = tmp_path / "r.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 3.5\n2 1 -1e3\n")
    row, col, *_ = bs.readCOO(str(p))
    assert row.tolist() == [0, 2, 2] and col.tolist() == [0, 1]
    p.write_text("%%MatrixMarket matrix coordinate complex general\n2 2 1\n2 2 1.0 2.0\n")
    row, col, *_ = bs.readCOO(str(p))
    assert row.tolist() == [0, 0, 1] and col.tolist() == [1]


@pytest.mark.parametrize("text,why", [
    ("%%NotMatrixMarket matrix coordinate pattern general\n1 1 0\n", "banner"),
    ("%%MatrixMarket matrix coordinate pattern\n1 1 0\n", "short banner"),
    ("%%MatrixMarket matrix array real general\n1 1\n", "dense"),
    ("%%MatrixMarket matrix coordinate pattern general\n2 2 2\n1 1\n", "premature eof"),
    ("%%MatrixMarket matrix coordinate pattern general\n2 2 1\n3 1\n", "index out of range"),
])
def test_readCOO_errors(bs, tmp_path, text, why):
    p = tmp_path / "bad.mtx"
    p.write_text(text)
    with pytest.raises(OSError):
        bs.readCOO(str(p))
    with pytest.raises(OSError):
        bs.readCOO(str(tmp_path / "missing.mtx"))


def test_readCOO_exit1_like_reference(bs, tmp_path):
    """The void readCOO exits with status 1 on failure (final/utils.c:54-61)."""
    code = ("import importlib,sys,ctypes;sys.path.insert(0,%r);b=importlib.import_module('binary-spgemm_b200');"
            "h=b.host();h.readCOO(b'/nonexistent.mtx',None,None,None,None,None)") % str(os.path.dirname(os.path.dirname(__file__)))
    r = subprocess.run(["python", "-c", code], capture_output=True)
    assert r.returncode == 1


def test_coo2csc_matches_oracle_and_is_stable(bs, oracle):
    rng = np.random.default_rng(3)
    for n in (1, 7, 100, 5000):
        nnz = int(rng.integers(0, 6 * n))
        r = rng.integers(0, n, nnz).astype(np.uint32)
        c = rng.integers(0, n, nnz).astype(np.uint32)
        a = bs.coo2csc(r, c, n)
        b = oracle.coo2csc(r, c, n)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all()
        # stability: inside each bucket the entries keep input order
        order = np.argsort(c, kind="stable")
        assert (a[0] == r[order]).all()
    a = bs.coo2csc(np.array([1, 2], np.uint32), np.array([2, 1], np.uint32), 2, 1)     # 1-based flag
    assert a[1].tolist() == [0, 1, 2] and a[0].tolist() == [1, 0]


def test_mm_banner_surface(bs, tmp_path):
    H = bs.host()
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    H.mm_read_banner.argtypes = [C.c_void_p, C.c_char_p]
    H.mm_read_mtx_crd_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    p = tmp_path / "b.mtx"
    p.write_text("%%MatrixMarket MATRIX Coordinate Integer Symmetric\n% c1\n% c2\n\n 5 6 7\n")
    f = libc.fopen(str(p).encode(), b"r")
    code = C.create_string_buffer(4)
    assert H.mm_read_banner(f, code) == 0 and code.raw == b"MCIS"
    M, N, nz = C.c_int(), C.c_int(), C.c_int()
    assert H.mm_read_mtx_crd_size(f, C.byref(M), C.byref(N), C.byref(nz)) == 0
    assert (M.value, N.value, nz.value) == (5, 6, 7)
    libc.fclose(f)
    for text, err in [("%%MatrixMarket vector coordinate real general\n", 15), ("%%Foo matrix coordinate real general\n", 14),
                      ("%%MatrixMarket matrix coordinate real\n", 12), ("%%MatrixMarket matrix coordinate quaternion general\n", 15)]:
        p.write_text(text)
        f = libc.fopen(str(p).encode(), b"r")
        assert H.mm_read_banner(f, code) == err          # MM_UNSUPPORTED_TYPE / MM_NO_HEADER / MM_PREMATURE_EOF
        libc.fclose(f)


def test_time_stats_like_reference(bs):
    """mean, lower median = sorted[(times-1)/2], fastest (final/SpGEMM_mpi_omp.c:330-333)."""
    H = bs.host()
    for vals in ([3.0, 1.0, 2.0, 4.0], [5.0], [2.0, 9.0, 4.0, 1.0, 7.0]):
        t = np.array(vals)
        mean, med, fast = C.c_double(), C.c_double(), C.c_double()
        H.bs_time_stats(t.ctypes.data, len(t), C.byref(mean), C.byref(med), C.byref(fast))
        s = sorted(vals)
        assert abs(mean.value - sum(vals) / len(vals)) < 1e-12 and med.value == s[(len(s) - 1) // 2] and fast.value == s[0]


def test_generators(bs, tmp_path):
    row, col = bs.gen_uniform(5000, 16, 1)
    assert row[0] == 0 and len(col) == row[-1] and col.max() < 5000
    lens = np.diff(row)
    assert lens.max() <= 16 and lens.min() >= 1
    for i in range(0, 5000, 500):
        assert (np.diff(col[row[i]:row[i + 1]]) > 0).all()
    r2, c2 = bs.gen_uniform(5000, 16, 1)
    assert (r2 == row).all() and (c2 == col).all()
    r3, c3 = bs.gen_uniform(5000, 16, 2)
    assert not (len(c3) == len(col) and (c3 == col).all())
    row, col = bs.gen_banded(100, 32)
    assert col[row[50]:row[51]].tolist() == list(range(34, 66))
    row, col = bs.gen_blockdiag(100, 32)
    assert col[row[40]:row[41]].tolist() == list(range(32, 64)) and col[row[99]:row[100]].tolist() == [96, 97, 98, 99]
    row, col = bs.gen_rmat(10, 8)
    assert len(row) == 1025 and col.max() < 1024
    # .mtx round trip through the transposing reader
    p = tmp_path / "g.mtx"
    bs.write_mtx(str(p), row, col)
    rr, cc, M, N, nnz = bs.readCOO(str(p))
    assert (rr == row).all() and (cc == col).all()


def test_readCOO_parallel_tokenizer_matches_the_oracle_reader(bs, oracle, tmp_path):
    """Files above 4 MB are tokenized by all host threads (host/utils.c parse_entries_parallel) when they are one entry per
    line; anything unusual (blank lines, two entries on a line, values) must give what the sequential tokenizer / the
    reference's fscanf loop give.  Compared with the oracle's restatement of readCOO (final/utils.c:47-81) and, when built,
    the compiled reference."""
    row, col = bs.gen_uniform(1 << 16, 8, 5)
    p = tmp_path / "big.mtx"
    bs.write_mtx(str(p), row, col)
    assert p.stat().st_size > (4 << 20)
    got = bs.readCOO(str(p))
    want = oracle.readCOO(p)
    assert got[2:] == want[2:] and (got[0] == want[0]).all() and (got[1] == want[1]).all()
    from oracle.oracle import Ref, have_ref
    if have_ref():
        ref = Ref().readCOO(p)
        assert (got[0] == ref[0]).all() and (got[1] == ref[1]).all()
    # the same entries in an irregular layout: CRLF, blank lines, two entries on one line -> sequential tokenizer, same arrays
    lines = p.read_text().splitlines()
    head = [l for l in lines if l.startswith("%")] + [next(l for l in lines if not l.startswith("%"))]
    body = lines[len(head):]
    odd = head + ["", body[0] + " " + body[1]] + body[2:1000] + ["   "] + [l + "\r" for l in body[1000:]]
    q = tmp_path / "odd.mtx"
    q.write_text("\n".join(odd) + "\n")
    got2 = bs.readCOO(str(q))
    assert (got2[0] == got[0]).all() and (got2[1] == got[1]).all()
    # real-valued file above the threshold: the value column is skipped by both tokenizers
    r = tmp_path / "real.mtx"
    r.write_text("\n".join([head[0].replace("pattern", "real")] + head[1:] + [l + " 1.5e0" for l in body]) + "\n")
    got3 = bs.readCOO(str(r))
    assert (got3[0] == got[0]).all() and (got3[1] == got[1]).all()
    # an index out of range somewhere in the middle is still an error
    bad = list(body); bad[len(bad) // 2] = "99999999 1"
    b = tmp_path / "bad.mtx"
    b.write_text("\n".join(head + bad) + "\n")
    with pytest.raises(OSError):
        bs.readCOO(str(b))


def test_work_balanced_row_blocks_are_contiguous_and_even(bs):
    """bench.py --split ip / BSPGEMM_SPLIT=ip: contiguous row blocks with equal intermediate products."""
    import importlib.util, sys, torch
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("bench_mod", Path(__file__).resolve().parents[1] / "bench.py")
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
    row, col = bs.gen_rmat(13, 16, 0.45, 0.22, 0.22, 1)
    n = len(row) - 1
    blen = np.diff(row).astype(np.int64)
    cs = np.concatenate([[0], np.cumsum(blen[col])])
    ip = cs[row[1:]] - cs[row[:-1]]
    for world in (2, 3, 8):
        b = bench.shard_bounds_ip(torch.from_numpy(row), torch.from_numpy(col), n, world)
        assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:])), b
        work = [int(ip[b[q]:b[q + 1]].sum()) + (b[q + 1] - b[q]) for q in range(world)]
        assert max(work) - min(work) <= 2 * int(ip.max()) + 2, (work, int(ip.max()))     # within one (largest) row of each other
        rows_split = [int(ip[(n * q) // world:(n * (q + 1)) // world].sum()) for q in range(world)]
        assert max(work) <= max(rows_split) + world                                      # never worse than equal rows
