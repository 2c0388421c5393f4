"""Parity of the CUDA path against the oracle — every call goes through the C ABI (include/bspgemm.h).
Bit-exact bar: identical Crow, identical ascending Ccol (integer/index work, no tolerance)."""
import numpy as np
import pytest

from helpers import assert_csr_contract, crc, gen_case, random_csr

pytestmark = pytest.mark.gpu


def _explain(got_col, got_row, want_col, want_row):
    want_row = np.asarray(want_row, np.int64)
    got_row = np.asarray(got_row, np.int64)
    if len(got_row) != len(want_row):
        return f"Crow length {len(got_row)} != {len(want_row)}"
    bad = np.nonzero(got_row != want_row)[0]
    if len(bad):
        i = int(bad[0])
        return f"Crow differs first at [{i}]: got {got_row[i]} want {want_row[i]} ({len(bad)} entries differ; nnz got {got_row[-1]} want {want_row[-1]})"
    bad = np.nonzero(got_col != want_col)[0]
    if len(bad):
        p = int(bad[0])
        r = int(np.searchsorted(want_row, p, side="right") - 1)
        s, e = int(want_row[r]), int(want_row[r + 1])
        return (f"Ccol differs first at [{p}] (row {r}, len {e - s}): got {got_col[s:e][:40].tolist()} "
                f"want {want_col[s:e][:40].tolist()} ({len(bad)} entries differ)")
    return ""


def check(bs, oracle, Acol, Arow, An, Bcol, Brow, Bn, Bm, mode=None):
    want_col, want_row = oracle.spgemm(Acol, Arow, An, Bcol, Brow, Bm)
    got_col, got_row = bs.spgemm_csr(Acol, Arow, An, Bcol, Brow, Bn, Bm)
    msg = _explain(got_col, got_row, want_col, want_row)
    assert not msg, msg
    return got_col, got_row


def dev_multiply(bs, mode, Acol, Arow, An, Bcol, Brow, Bn, Bm, i64=False):
    import torch
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
    dAc, dAr, dBc, dBr = t(Acol), t(Arow), t(Bcol), t(Brow)
    dCr = torch.full((An + 1,), -7, dtype=torch.int64 if i64 else torch.int32, device=dev)
    h = bs.DeviceSpGEMM(0, mode)
    ptr, nnz = h.multiply(dAc, dAr, An, len(Acol), dBc, dBr, Bn, Bm, len(Bcol), dCr, crow_is_i64=i64)
    torch.cuda.synchronize()
    col = bs.device_view(ptr, nnz, 0).cpu().numpy().copy()
    st = h.stats()
    h.close()
    return col, dCr.cpu().numpy(), st


def test_fixture_golden(gpu_ctx, fixture_npz):
    """Config 1: the reference's validity_test.mtx, C = A·A; absolute crc32s from SURVEY.md §4."""
    f = fixture_npz
    M = int(f["M"])
    Ccol, Crow = gpu_ctx.spgemm_csr(f["Acol"], f["Arow"], M, f["Acol"], f["Arow"], M, M)
    assert crc(Crow) == "62b292d7" and crc(Ccol) == "77f3757b"
    assert (Crow == f["Crow"]).all() and (Ccol == f["Ccol"]).all()
    assert gpu_ctx.intermediate_products(f["Acol"], f["Arow"], M, f["Arow"], M) == 12502


def test_kats(gpu_ctx, kats):
    for k in kats:
        An, Bn = len(k["Arow"]) - 1, len(k["Brow"]) - 1
        Ccol, Crow = gpu_ctx.spgemm_csr(k["Acol"], k["Arow"], An, k["Bcol"], k["Brow"], Bn, k["Bm"])
        assert Crow.tolist() == k["Crow"] and Ccol.tolist() == k["Ccol"], k["name"]


def test_seeded_cases_golden(gpu_ctx, seeded_cases):
    """Committed checksums produced by the compiled reference (tests/golden/make_golden.py)."""
    for c in seeded_cases:
        row, col = gen_case(gpu_ctx, c)
        n = c["n"]
        Ccol, Crow = gpu_ctx.spgemm_csr(col, row, n, col, row, n, n)
        assert int(Crow[-1]) == c["nnzC"], c["name"]
        assert crc(Crow) == c["Crow_crc"] and crc(Ccol) == c["Ccol_crc"], c["name"]
        assert_csr_contract(Ccol, Crow, n)


@pytest.mark.parametrize("mode", ["fused", "twophase"])
def test_device_operator_modes(bs, oracle, seeded_cases, mode):
    m = bs.MODE_FUSED if mode == "fused" else bs.MODE_TWOPHASE
    for c in seeded_cases:
        row, col = gen_case(bs, c)
        n = c["n"]
        want_col, want_row = oracle.spgemm(col, row, n, col, row, n)
        got_col, got_row, st = dev_multiply(bs, m, col, row, n, col, row, n, n)
        msg = _explain(got_col, got_row, want_col, want_row)
        assert not msg, f"{c['name']} [{mode}] {msg}"
        assert st["mode"] == m and st["nnz"] == len(want_col)
        assert st["ip"] == oracle.intermediate_products(col, row, n, row)


def test_forced_estimate_path_matches(bs, oracle, seeded_cases, monkeypatch):
    """Small-degree inputs normally skip the work-estimation pass (maxlen(A)*maxlen(B) bounds every row);
    force it so that estimate -> fused is covered on the same inputs."""
    monkeypatch.setenv("BSPGEMM_FORCE_ESTIMATE", "1")
    for c in seeded_cases[:4]:
        row, col = gen_case(bs, c)
        n = c["n"]
        want_col, want_row = oracle.spgemm(col, row, n, col, row, n)
        got_col, got_row, st = dev_multiply(bs, bs.MODE_FUSED, col, row, n, col, row, n, n)
        msg = _explain(got_col, got_row, want_col, want_row)
        assert not msg, f"{c['name']} {msg}"
        assert st["ip"] == oracle.intermediate_products(col, row, n, row)


def test_small_capacity_forces_cta_bins(bs, oracle, monkeypatch):
    """BSPGEMM_CAP_S=32: most rows of a d=8 matrix (IP ~ 64) leave the warp bin -> M symbolic + fused + M numeric."""
    monkeypatch.setenv("BSPGEMM_CAP_S", "32")
    row, col = bs.gen_uniform(20000, 8, 11)
    n = 20000
    want_col, want_row = oracle.spgemm(col, row, n, col, row, n)
    for mode in (bs.MODE_FUSED, bs.MODE_TWOPHASE):
        got_col, got_row, st = dev_multiply(bs, mode, col, row, n, col, row, n, n)
        msg = _explain(got_col, got_row, want_col, want_row)
        assert not msg, f"mode {mode}: {msg}"
        assert st["cap_s"] == 32 and st["rows_m"] > 0


def test_random_rectangular_unsorted_duplicates(gpu_ctx, oracle):
    """A(n x k) · B(k x m), unsorted rows and repeated columns on input (legal for the reference)."""
    rng = np.random.default_rng(17)
    for trial in range(40):
        n, k, m = (int(x) for x in rng.integers(1, 700, 3))
        Arow, Acol = random_csr(rng, n, k, rng.uniform(0, 8), sort=trial % 2 == 0, dups=trial % 3 == 0)
        Brow, Bcol = random_csr(rng, k, m, rng.uniform(0, 8), sort=trial % 2 == 1, dups=trial % 3 == 1)
        check(gpu_ctx, oracle, Acol, Arow, n, Bcol, Brow, k, m)


def test_edge_cases(gpu_ctx, oracle):
    z = np.zeros(0, np.int32)
    # no rows at all
    Ccol, Crow = gpu_ctx.spgemm_csr(z, [0], 0, z, [0], 0, 0)
    assert Crow.tolist() == [0] and len(Ccol) == 0
    # rows but no entries
    Ccol, Crow = gpu_ctx.spgemm_csr(z, [0] * 6, 5, z, [0] * 4, 3, 9)
    assert Crow.tolist() == [0] * 6 and len(Ccol) == 0
    # A nonzeros that only select empty B rows
    check(gpu_ctx, oracle, [0, 1, 2, 1], [0, 2, 4], 2, [5], [0, 0, 0, 0, 1], 4, 8)
    # one dense row (IP = n*n/…): every column present
    n = 300
    Arow = np.array([0, n], np.int32)
    Acol = np.arange(n, dtype=np.int32)
    Brow = (np.arange(n + 1) * n).astype(np.int32)
    Bcol = np.tile(np.arange(n, dtype=np.int32), n)
    Ccol, Crow = check(gpu_ctx, oracle, Acol, Arow, 1, Bcol, Brow, n, n)
    assert Ccol.tolist() == list(range(n))
    # single-entry matrices, last column
    check(gpu_ctx, oracle, [0], [0, 1], 1, [6], [0, 1], 1, 7)
    # ragged: very long A row pointing at mostly empty B rows
    k = 5000
    Arow = np.array([0, k, k, k + 3], np.int32)
    Acol = np.concatenate([np.arange(k), [1, 2, 3]]).astype(np.int32)
    blen = np.zeros(k, np.int64); blen[::97] = 3
    Brow = np.concatenate([[0], np.cumsum(blen)]).astype(np.int32)
    Bcol = (np.arange(Brow[-1]) * 7 % 1000).astype(np.int32)
    check(gpu_ctx, oracle, Acol, Arow, 3, Bcol, Brow, k, 1000)


def test_every_bin_is_exercised(bs, oracle):
    """Rows for the warp bin (hash + bitmap), both CTA bins and the global-bitmap bin in one matrix."""
    rng = np.random.default_rng(23)
    Bm = 1 << 20
    k = 4096
    blen = rng.integers(0, 40, k)
    blen[:64] = 3000                                # long B rows -> huge IP for rows that select them
    Brow = np.concatenate([[0], np.cumsum(blen)]).astype(np.int32)
    Bcol = np.concatenate([np.sort(rng.choice(Bm, l, replace=False)) for l in blen]).astype(np.int32)
    rows = []
    rows.append(rng.choice(np.arange(64, k), 10, replace=False))          # S bin, wide span -> ordered table
    rows.append(np.arange(64, 64 + 300))                                   # IP ~ 6000 -> CTA bin (M2)
    rows.append(np.arange(64, 64 + 60))                                    # IP ~ 1200 -> CTA bin (M1)
    rows.append(np.arange(0, 20))                                          # IP = 60000 -> global bitmap
    rows.append(np.arange(0, 64))                                          # IP = 192000 -> global bitmap
    rows.append(np.zeros(0, np.int64))
    for _ in range(200):
        rows.append(rng.choice(np.arange(64, k), int(rng.integers(0, 12)), replace=False))
    Arow = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
    Acol = np.concatenate(rows).astype(np.int32)
    An = len(rows)
    want_col, want_row = oracle.spgemm(Acol, Arow, An, Bcol, Brow, Bm)
    for mode in (bs.MODE_FUSED, bs.MODE_TWOPHASE):
        got_col, got_row, st = dev_multiply(bs, mode, Acol, Arow, An, Bcol, Brow, k, Bm)
        msg = _explain(got_col, got_row, want_col, want_row)
        assert not msg, f"mode {mode}: {msg}"
        assert st["rows_m"] > 0 and st["rows_l"] > 0 and st["rows_s"] > 0


def test_narrow_span_rows_use_bitmap_and_match(gpu_ctx, oracle, monkeypatch):
    """Clustered columns (banded / block-diagonal): the [lo,hi] bitmap path inside the warp and CTA bins."""
    monkeypatch.setenv("BSPGEMM_NO_BAND", "1")       # not the run/bitmap kernel of band.cuh (tested on its own)
    for gen, args in ((gpu_ctx.gen_banded, (3000, 32)), (gpu_ctx.gen_blockdiag, (3000, 32)), (gpu_ctx.gen_banded, (500, 8))):
        row, col = gen(*args)
        n = len(row) - 1
        check(gpu_ctx, oracle, col, row, n, col, row, n, n)


def test_i64_row_pointers_and_slices(gpu_ctx, oracle, bs):
    row, col = bs.gen_uniform(20000, 8, 9)
    n = 20000
    want_col, want_row = oracle.spgemm(col, row, n, col, row, n)
    Ccol, Crow = bs.spgemm_csr(col, row, n, col, row, n, n, i64=True)
    assert Crow.dtype == np.int64 and (Crow == want_row).all() and (Ccol == want_col).all()
    # SpGEMM_bigslice replacement: slice-relative row pointers (final/SpGEMM_mpi_omp.c:20,26,54)
    s, e = 3333, 17001
    Scol, Srow = bs.spgemm_csr_slice(col, row, n, col, row, n, n, s, e)
    assert (Srow == want_row[s:e + 1] - want_row[s]).all()
    assert (Scol == want_col[want_row[s]:want_row[e]]).all()
    # shifted Arow pointer with absolute offsets, like &Arow[rank*tasksize] (:171)
    Pcol, Prow = bs.spgemm_csr(col, row[s:], e - s, col, row, n, n)
    assert (Prow == Srow).all() and (Pcol == Scol).all()


def test_caller_allocated_output_and_legacy_signature(gpu_ctx, oracle, bs):
    row, col = bs.gen_uniform(3000, 8, 4)
    n = 3000
    want_col, want_row = oracle.spgemm(col, row, n, col, row, n)
    buf = np.empty(len(want_col) + 10, np.int32)
    nnz, Crow = bs.spgemm_csr_into(col, row, n, col, row, n, n, buf)
    assert nnz == len(want_col) and (buf[:nnz] == want_col).all() and (Crow == want_row).all()
    with pytest.raises(bs.BSpGEMMError) as ei:
        bs.spgemm_csr_into(col, row, n, col, row, n, n, np.empty(5, np.int32))
    assert ei.value.status == bs.ERR_CAPACITY
    Lcol, Lrow = bs.SpGEMM_mpi(col, row, n, col, row, n, 375)      # reference argument list, no Bn
    assert (Lrow == want_row).all() and (Lcol == want_col).all()


def test_bad_arguments_are_reported_not_computed(gpu_ctx, bs):
    with pytest.raises(bs.BSpGEMMError) as ei:
        bs.spgemm_csr([0, 5], [0, 2], 1, [0], [0, 1, 1], 2, 4)      # A column 5 outside Bn=2
    assert ei.value.status == bs.ERR_BADARG
    with pytest.raises(bs.BSpGEMMError) as ei:
        bs.spgemm_csr([0], [0, 1], 1, [9], [0, 1], 1, 4)            # B column 9 outside Bm=4
    assert ei.value.status == bs.ERR_BADARG
    # the context stays usable afterwards
    Ccol, Crow = bs.spgemm_csr([0], [0, 1], 1, [3], [0, 1], 1, 4)
    assert Ccol.tolist() == [3] and Crow.tolist() == [0, 1]


def test_repeated_calls_reuse_the_context(gpu_ctx, oracle, bs):
    """The driver calls the operator `times` times (final/SpGEMM_mpi_omp.c:318-328)."""
    row, col = bs.gen_uniform(10000, 8, 2)
    want_col, want_row = oracle.spgemm(col, row, 10000, col, row, 10000)
    for _ in range(5):
        Ccol, Crow = bs.spgemm_csr(col, row, 10000, col, row, 10000, 10000)
        assert (Crow == want_row).all() and (Ccol == want_col).all()
    row2, col2 = bs.gen_banded(777, 16)
    Ccol, Crow = bs.spgemm_csr(col2, row2, 777, col2, row2, 777, 777)
    w2c, w2r = oracle.spgemm(col2, row2, 777, col2, row2, 777)
    assert (Crow == w2r).all() and (Ccol == w2c).all()


def test_config2_full_size_bit_exact(bs, oracle):
    """BASELINE config 2: uniform random n=2^20, d=8, C = A·A — full bit-exact compare with the oracle."""
    n = 1 << 20
    row, col = bs.gen_uniform(n, 8, 1)
    want_col, want_row = oracle.spgemm(col, row, n, col, row, n)
    for mode in (bs.MODE_FUSED, bs.MODE_TWOPHASE):
        got_col, got_row, st = dev_multiply(bs, mode, col, row, n, col, row, n, n)
        assert st["ip"] == oracle.intermediate_products(col, row, n, row)
        msg = _explain(got_col, got_row, want_col, want_row)
        assert not msg, msg


def test_config3_full_size_properties(bs, oracle):
    """BASELINE config 3 (the bench workload): n=2^22, d=16.  Size-independent properties on the full output
    (row-pointer monotonicity, strict ascent inside rows, nnz <= IP, fused == two-phase checksums) plus a
    bit-exact compare of a contiguous row block and of scattered rows against the oracle."""
    import torch
    n = 1 << 22
    row, col = bs.gen_uniform(n, 16, 1)
    dev = torch.device("cuda:0")
    dAr, dAc = torch.from_numpy(row).to(dev), torch.from_numpy(col).to(dev)
    sums = []
    for mode in (bs.MODE_FUSED, bs.MODE_TWOPHASE):
        h = bs.DeviceSpGEMM(0, mode)
        dCr = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        ptr, nnz = h.multiply(dAc, dAr, n, len(col), dAc, dAr, n, n, len(col), dCr)
        torch.cuda.synchronize()
        st = h.stats()
        C = bs.device_view(ptr, nnz, 0)
        Cr = dCr.to(torch.int64)
        assert int(Cr[0]) == 0 and int(Cr[-1]) == nnz and bool((Cr[1:] >= Cr[:-1]).all())
        assert nnz <= st["ip"] and st["ip"] == int((torch.from_numpy(np.diff(row).astype(np.int64))[torch.from_numpy(col).long()]).sum())
        # strictly ascending except across row boundaries
        d = C[1:] > C[:-1]
        starts = Cr[1:-1]
        starts = starts[(starts > 0) & (starts < nnz)]
        d[starts - 1] = True
        assert bool(d.all())
        assert int(C.min()) >= 0 and int(C.max()) < n
        sums.append((nnz, int(C.to(torch.int64).sum()), int((C.to(torch.int64) * (torch.arange(nnz, device=dev) % 1000003)).sum()), int(Cr.sum())))
        if mode == bs.MODE_FUSED:
            # bit-exact block + scattered rows vs the oracle
            r0, r1 = 1234567, 1234567 + 50000
            wc, wr = oracle.spgemm(col, row[r0:], r1 - r0, col, row, n)
            gr = Cr[r0:r1 + 1].cpu().numpy()
            assert (gr - gr[0] == wr).all()
            assert (C[gr[0]:gr[-1]].cpu().numpy() == wc).all()
            for r in (0, 1, n // 2, n - 1):
                wc, wr = oracle.spgemm(col, row[r:], 1, col, row, n)
                assert (C[int(Cr[r]):int(Cr[r + 1])].cpu().numpy() == wc).all()
        del C
        h.close()
    assert sums[0] == sums[1]


def test_ell_fast_path_variants(bs, oracle, monkeypatch):
    """The ELL fast path (fused_ell.cuh): every ELL width W and tile height R, ragged / empty / long A rows,
    unsorted B rows, repeated A columns (every key duplicated), 64-bit row pointers."""
    monkeypatch.setenv("BSPGEMM_FORCE_ELL", "1")     # skip the padding-waste rule: max_len(B) <= 32 is enough
    rng = np.random.default_rng(31)
    cases = []
    for d in (3, 6, 12, 24):                       # W = 4, 8, 16, 32
        row, col = bs.gen_uniform(30000, d, 5 + d)
        cases.append((f"uniform d={d}", col, row, 30000, col, row, 30000, 30000))
    # A: ragged row lengths 0..40 (tiles with E > 64, empty rows), B: d<=16 unsorted
    n, m = 20000, 1 << 18
    Arow, Acol = random_csr(rng, n, n, 14.0, sort=False, dups=True)
    Brow, Bcol = random_csr(rng, n, m, 10.0, sort=True, dups=False)
    Bcol = Bcol.copy()
    for i in range(n):                               # shuffle inside every B row: nothing guarantees sorted input rows
        rng.shuffle(Bcol[Brow[i]:Brow[i + 1]])
    cases.append(("ragged A, unsorted B", Acol, Arow, n, Bcol, Brow, n, m))
    # short A rows (R = 8), tiny B rows
    Arow, Acol = random_csr(rng, n, n, 2.0, sort=True, dups=False)
    Brow, Bcol = random_csr(rng, n, m, 2.5, sort=True, dups=False)
    cases.append(("short rows", Acol, Arow, n, Bcol, Brow, n, m))
    seen = set()
    for kernel in ("hash", "sort", "sort-sync", "sort-async"):
        # hash: ordered-table kernel (fused_ell.cuh, variant 1); sort: register sorting network (fused_sort.cuh, variant 2),
        # as dispatched / k_fused_sort (register prefetch) for every geometry / k_fused_sort_async (cp.async) for every geometry
        for v in ("BSPGEMM_NO_SORT", "BSPGEMM_FORCE_SORT", "BSPGEMM_SORT_SYNC", "BSPGEMM_SORT_ASYNC"):
            monkeypatch.delenv(v, raising=False)
        monkeypatch.setenv("BSPGEMM_NO_SORT" if kernel == "hash" else "BSPGEMM_FORCE_SORT", "1")
        if kernel == "sort-sync":
            monkeypatch.setenv("BSPGEMM_SORT_SYNC", "1")
        if kernel == "sort-async":
            monkeypatch.setenv("BSPGEMM_SORT_ASYNC", "1")
        for name, Acol, Arow, An, Bcol, Brow, Bn, Bm in cases:
            want_col, want_row = oracle.spgemm(Acol, Arow, An, Bcol, Brow, Bm)
            for i64 in (False, True):
                got_col, got_row, st = dev_multiply(bs, bs.MODE_AUTO, Acol, Arow, An, Bcol, Brow, Bn, Bm, i64=i64)
                msg = _explain(got_col, got_row, want_col, want_row)
                assert not msg, f"{name} [{kernel}] i64={i64}: {msg}"
                assert st["ip"] == oracle.intermediate_products(Acol, Arow, An, Brow), name
                if np.diff(Brow).max() <= 32:
                    la = 4
                    while la < np.diff(Arow).max():
                        la *= 2
                    sortable = la <= 32 and la * st["group"] <= 1024
                    want_variant = 2 if (kernel != "hash" and sortable) else 1
                    assert st["variant"] == want_variant, f"{name} [{kernel}]: stats {st}"
                    seen.add((st["variant"], st["group"], st["rows_per_tile"]))
    assert {w for v, w, _ in seen if v == 1} == {4, 8, 16, 32}, seen
    assert {w for v, w, _ in seen if v == 2} >= {4, 8, 16}, seen
    assert len({r for v, _, r in seen if v == 1}) >= 3, seen


@pytest.mark.parametrize("kernel", ["hash", "sort"])
def test_ell_clustered_columns_spill_and_rebuild(bs, oracle, monkeypatch, kernel):
    """Columns of B concentrated in a sliver of a wide [0,Bm): the global monotone slot map sends every key of a
    row to a handful of slots, chains run past the 32 spare slots, the row is rebuilt by the exact path."""
    monkeypatch.setenv("BSPGEMM_NO_SORT" if kernel == "hash" else "BSPGEMM_FORCE_SORT", "1")
    monkeypatch.setenv("BSPGEMM_FORCE_ELL", "1")     # the span probe would route clustered rows to the bitmap kernels
    rng = np.random.default_rng(37)
    n, Bm = 6000, 1 << 22
    Arow, Acol = random_csr(rng, n, n, 12.0, sort=True, dups=False)
    blen = rng.integers(8, 17, n)
    Brow = np.concatenate([[0], np.cumsum(blen)]).astype(np.int32)
    base = 3_000_000
    Bcol = np.concatenate([np.sort(rng.choice(300, l, replace=False)) + base for l in blen]).astype(np.int32)
    want_col, want_row = oracle.spgemm(Acol, Arow, n, Bcol, Brow, Bm)
    got_col, got_row, st = dev_multiply(bs, bs.MODE_AUTO, Acol, Arow, n, Bcol, Brow, n, Bm)
    msg = _explain(got_col, got_row, want_col, want_row)
    assert not msg, msg
    assert st["variant"] == (1 if kernel == "hash" else 2)
    # two clusters far apart + a few uniform columns: partial spills
    Bcol2 = Bcol.copy()
    sel = rng.random(len(Bcol2)) < 0.3
    Bcol2[sel] = rng.integers(0, Bm, int(sel.sum()))
    want_col, want_row = oracle.spgemm(Acol, Arow, n, Bcol2, Brow, Bm)
    got_col, got_row, st = dev_multiply(bs, bs.MODE_AUTO, Acol, Arow, n, Bcol2, Brow, n, Bm)
    msg = _explain(got_col, got_row, want_col, want_row)
    assert not msg, msg


def test_big_rows_two_pass_when_staging_is_off(bs, oracle, monkeypatch):
    """BSPGEMM_NO_STAGE: M/L rows counted, then filled (the route taken when the staging arena does not fit in memory)."""
    monkeypatch.setenv("BSPGEMM_NO_STAGE", "1")
    test_every_bin_is_exercised(bs, oracle)
    row, col = bs.gen_rmat(13, 16, 0.57, 0.19, 0.19, 7)
    n = len(row) - 1
    want_col, want_row = oracle.spgemm(col, row, n, col, row, n)
    got_col, got_row, st = dev_multiply(bs, bs.MODE_FUSED, col, row, n, col, row, n, n)
    msg = _explain(got_col, got_row, want_col, want_row)
    assert not msg, msg


def test_power_law_rows_sort_and_window_bins(bs, oracle):
    """R-MAT (no vertex permutation): the candidates of every row pile up on the hub columns.  Small rows go through the
    register sort of the warp bin, big rows through the windowed shared-memory bitmap (rows_window.cuh), hub rows of B
    through the warp-per-long-row loop."""
    row, col = bs.gen_rmat(15, 16, 0.45, 0.22, 0.22, 3)
    n = len(row) - 1
    want_col, want_row = oracle.spgemm(col, row, n, col, row, n)
    for mode in (bs.MODE_FUSED, bs.MODE_TWOPHASE):
        got_col, got_row, st = dev_multiply(bs, mode, col, row, n, col, row, n, n, i64=True)
        msg = _explain(got_col, got_row, want_col, want_row)
        assert not msg, f"mode {mode}: {msg}"
        assert st["rows_m"] > 0 and st["rows_l"] > 0 and st["rows_s"] > 0
    # a steeper one (Graph500 parameters): longer hub rows
    row, col = bs.gen_rmat(13, 16, 0.57, 0.19, 0.19, 5)
    n = len(row) - 1
    want_col, want_row = oracle.spgemm(col, row, n, col, row, n)
    got_col, got_row, st = dev_multiply(bs, bs.MODE_FUSED, col, row, n, col, row, n, n, i64=True)
    msg = _explain(got_col, got_row, want_col, want_row)
    assert not msg, msg


def test_window_bins_many_windows_unsorted_rows(bs, oracle, monkeypatch):
    """Big rows whose columns span several bitmap windows (Bm = 6M columns = 4 windows), with gaps, with B rows that are
    NOT ascending (the first-window guess is wrong: restart path) and with repeated columns."""
    rng = np.random.default_rng(29)
    Bm = 6_000_000
    k = 3000
    blen = rng.integers(1, 30, k)
    blen[:40] = 2500                                  # long B rows (warp-per-row loop)
    Brow = np.concatenate([[0], np.cumsum(blen)]).astype(np.int32)
    parts = []
    for i, l in enumerate(blen):
        if i % 3 == 0:   c = rng.integers(0, Bm, l)                                   # anywhere, unsorted, may repeat
        elif i % 3 == 1: c = rng.integers(4_000_000, 4_000_000 + 5000, l)             # a cluster in window 2
        else:            c = np.sort(rng.integers(5_900_000, Bm, l))                  # the last columns
        parts.append(c)
    Bcol = np.concatenate(parts).astype(np.int32)
    rows = [np.arange(0, 40), np.arange(1, 2000, 3), np.arange(2, 2000, 3), np.arange(40, 1500), np.arange(0, k),
            rng.choice(k, 200, replace=False), np.arange(41, 3000, 3),
            np.arange(42, 3000, 3)[:600],      # ~9 K scattered products: the compressed single pass of rows_bm.cuh (piece slots)
            np.arange(42, 3000, 3)[:960]]      # ~14 K scattered products over 6 M columns: more than 11264 pieces -> back to windows
    for _ in range(50):
        rows.append(rng.choice(np.arange(40, k), int(rng.integers(0, 10)), replace=False))
    Arow = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
    Acol = np.concatenate(rows).astype(np.int32)
    An = len(rows)
    want_col, want_row = oracle.spgemm(Acol, Arow, An, Bcol, Brow, Bm)
    for cap in (None, "32"):
        if cap: monkeypatch.setenv("BSPGEMM_CAP_S", cap)
        for mode in (bs.MODE_FUSED, bs.MODE_TWOPHASE):
            got_col, got_row, st = dev_multiply(bs, mode, Acol, Arow, An, Bcol, Brow, k, Bm)
            msg = _explain(got_col, got_row, want_col, want_row)
            assert not msg, f"cap {cap} mode {mode}: {msg}"
            assert st["rows_m"] > 0 and st["rows_l"] > 0


def test_wide_matrices_keep_table_and_global_bitmap_bins(bs, oracle, monkeypatch):
    """BSPGEMM_NO_WINDOW: the global-bitmap kernel of the L bin (the route of matrices with more than WIN_MAX_WINDOWS
    windows of columns) still matches."""
    monkeypatch.setenv("BSPGEMM_NO_WINDOW", "1")
    monkeypatch.setenv("BSPGEMM_NO_BM", "1")
    test_every_bin_is_exercised(bs, oracle)


def test_round1_big_row_kernels_still_match(bs, oracle, monkeypatch):
    """BSPGEMM_NO_BM: the CTA-wide sort (rows_sort.cuh, 2049..16384 products) and the group-per-B-row window kernel
    (rows_window.cuh) — the route of matrices wider than BM_MAX_WINDOWS windows for the sort — still match."""
    monkeypatch.setenv("BSPGEMM_NO_BM", "1")
    test_power_law_rows_sort_and_window_bins(bs, oracle)
    test_window_bins_many_windows_unsorted_rows(bs, oracle, monkeypatch)


def test_big_rows_dense_and_sparse_words(bs, oracle, monkeypatch):
    """rows_bm.cuh: rows whose columns are long runs (dense bitmap words) next to scattered ones, A rows longer than one
    chunk of 1024 entries, empty B rows in between, two windows."""
    rng = np.random.default_rng(41)
    Bm, k = 2_500_000, 4000
    parts, blen = [], []
    for i in range(k):
        t = i % 4
        if t == 0:   s0 = int(rng.integers(0, Bm - 4000)); c = np.arange(s0, s0 + int(rng.integers(500, 4000)))      # a run
        elif t == 1: c = np.unique(rng.integers(0, Bm, int(rng.integers(1, 200))))                                   # scattered
        elif t == 2: c = np.zeros(0, np.int64)                                                                       # empty
        else:        s0 = int(rng.integers(0, 30000)); c = np.arange(s0, s0 + 64 * int(rng.integers(1, 40)), 2)      # every other column, low range
        parts.append(c); blen.append(len(c))
    Brow = np.concatenate([[0], np.cumsum(blen)]).astype(np.int32)
    Bcol = np.concatenate(parts).astype(np.int32)
    rows = [np.arange(0, k), rng.choice(k, 2500, replace=False), np.arange(0, k, 4)[:300], np.arange(1, k, 4)[:900], np.arange(3, k, 4)]
    for _ in range(200):
        rows.append(rng.choice(k, int(rng.integers(1, 120)), replace=False))
    Arow = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
    Acol = np.concatenate(rows).astype(np.int32)
    An = len(rows)
    want_col, want_row = oracle.spgemm(Acol, Arow, An, Bcol, Brow, Bm)
    for mode in (bs.MODE_FUSED, bs.MODE_TWOPHASE):
        got_col, got_row, st = dev_multiply(bs, mode, Acol, Arow, An, Bcol, Brow, k, Bm, i64=True)
        msg = _explain(got_col, got_row, want_col, want_row)
        assert not msg, f"mode {mode}: {msg}"
        assert st["rows_m"] > 0 and st["rows_l"] > 0


def _band_csr(n, m, below, above):
    """Row i holds the run of columns [i-below, i+above] clipped to [0,m)."""
    lo = np.clip(np.arange(n) - below, 0, m)
    hi = np.clip(np.arange(n) + above + 1, 0, m)
    hi = np.maximum(hi, lo)
    row = np.concatenate([[0], np.cumsum(hi - lo)]).astype(np.int32)
    col = np.concatenate([np.arange(a, b) for a, b in zip(lo, hi)]).astype(np.int32) if row[-1] else np.zeros(0, np.int32)
    return row, col


def test_band_kernel_runs_and_register_bitmap(bs, oracle):
    """BASELINE config 5 shapes (banded, block-diagonal): B rows as (first,len) runs, output rows as 128-bit bitmaps."""
    cases = [("banded d=32", *bs.gen_banded(5000, 32)), ("blockdiag d=32", *bs.gen_blockdiag(5000, 32)), ("banded d=8", *bs.gen_banded(777, 8))]
    for name, row, col in cases:
        n = len(row) - 1
        want_col, want_row = oracle.spgemm(col, row, n, col, row, n)
        got_col, got_row, st = dev_multiply(bs, bs.MODE_AUTO, col, row, n, col, row, n, n)
        msg = _explain(got_col, got_row, want_col, want_row)
        assert not msg, f"{name}: {msg}"
        assert st["variant"] == 3, f"{name}: variant {st['variant']}"
        assert st["ip"] == oracle.intermediate_products(col, row, n, row)
    # rectangular, A rows longer than a warp (40 runs per row), empty rows of A and of B
    Arow, Acol = _band_csr(3000, 2500, 20, 19)
    Brow, Bcol = _band_csr(2500, 4000, 10, 9)
    keep = np.ones(len(Arow) - 1, bool); keep[100:140] = False; keep[2999] = False          # empty rows of A
    lens = np.diff(Arow) * keep
    Acol = np.concatenate([Acol[Arow[i]:Arow[i + 1]] for i in np.nonzero(keep)[0]]).astype(np.int32)
    Arow = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    want_col, want_row = oracle.spgemm(Acol, Arow, 3000, Bcol, Brow, 4000)
    got_col, got_row, st = dev_multiply(bs, bs.MODE_AUTO, Acol, Arow, 3000, Bcol, Brow, 2500, 4000)
    msg = _explain(got_col, got_row, want_col, want_row)
    assert not msg, msg
    assert st["variant"] == 3


def test_band_kernel_falls_back_when_a_row_is_not_a_run(bs, oracle):
    """The probe samples rows; a defect elsewhere (a B row with a gap, an output row wider than the bitmap) must be caught
    by the kernels themselves and the product redone by the general path — same result."""
    n = 100_000
    Arow, Acol = _band_csr(n, n, 16, 15)
    # (1) B row 20 (not visited by the sampled rows 0, 48, 97, ...) has a gap
    Brow, Bcol = _band_csr(n, n, 16, 15)
    Bcol = Bcol.copy(); Bcol[Brow[20] + 3] += 0      # still a run
    want_col, want_row = oracle.spgemm(Acol, Arow, n, Bcol, Brow, n)
    got_col, got_row, st = dev_multiply(bs, bs.MODE_AUTO, Acol, Arow, n, Bcol, Brow, n, n)
    assert not _explain(got_col, got_row, want_col, want_row) and st["variant"] == 3
    Bcol[Brow[20 + 1] - 1] += 40                     # last entry of row 20 jumps: not a run any more
    want_col, want_row = oracle.spgemm(Acol, Arow, n, Bcol, Brow, n)
    got_col, got_row, st = dev_multiply(bs, bs.MODE_AUTO, Acol, Arow, n, Bcol, Brow, n, n)
    msg = _explain(got_col, got_row, want_col, want_row)
    assert not msg, msg
    assert st["variant"] != 3
    # (2) every B row is a run, but row 7 of A also selects a far-away B row: its output spans more than 128 columns
    Brow, Bcol = _band_csr(n, n, 16, 15)
    A2col = Acol.copy(); A2col[Arow[7 + 1] - 1] = 5000
    want_col, want_row = oracle.spgemm(A2col, Arow, n, Bcol, Brow, n)
    got_col, got_row, st = dev_multiply(bs, bs.MODE_AUTO, A2col, Arow, n, Bcol, Brow, n, n)
    msg = _explain(got_col, got_row, want_col, want_row)
    assert not msg, msg
    assert st["variant"] != 3


def test_gpu_coo2csc_is_the_stable_counting_sort(bs, oracle):
    """bspgemm_coo2csc (csrc/coo2csc.cuh) against the host coo2csc and the oracle's restatement of
    final/coo2csc.c:22-64: identical pointers, identical (stable: input order inside a column) index array.  Unsorted entries,
    repeated coordinates, both index bases, n from one radix pass (<= 256) to three, ragged tails of the 4096-entry chunks."""
    rng = np.random.default_rng(77)
    cases = [(1, 0), (1, 1), (7, 50), (200, 5000), (256, 4096), (257, 4097), (5000, 123457), (1 << 16, 300001),
             ((1 << 20) + 3, 2_000_003), (1 << 22, 6_000_000)]
    for n, nnz in cases:
        for one in (0, 1):
            I = rng.integers(0, n, nnz, dtype=np.uint32) + one
            J = rng.integers(0, n, nnz, dtype=np.uint32) + one
            if nnz > 10:                                   # a heavy column and exact duplicates
                J[rng.integers(0, nnz, nnz // 7)] = J[0]
                I[1], J[1] = I[0], J[0]
            want_row, want_col = bs.coo2csc(I, J, n, one)
            got_row, got_col = bs.coo2csc_gpu(I, J, n, one)
            assert (got_col == want_col).all(), (n, nnz, one)
            assert (got_row == want_row).all(), (n, nnz, one)
            if nnz <= 200000:
                o_row, o_col = oracle.coo2csc(I, J, n, one)
                assert (got_col == o_col).all() and (got_row == o_row).all(), (n, nnz, one)
    # a key outside [0,n) is an error, not a silent drop or an out-of-bounds write
    with pytest.raises(bs.BSpGEMMError) as e:
        bs.coo2csc_gpu(np.array([0, 1], np.uint32), np.array([0, 9], np.uint32), 4, 0)
    assert e.value.status == bs.ERR_BADARG


def test_gpu_coo2csc_feeds_the_product(bs, oracle, tmp_path):
    """Matrix Market text -> host tokenizer -> GPU coo2csc (readCOO_convert + bspgemm_coo2csc) gives the arrays readCOO
    gives, and the GPU product of them matches the oracle."""
    row, col = bs.gen_uniform(3000, 6, 9)
    path = tmp_path / "m.mtx"
    bs.write_mtx(str(path), row, col)
    want = bs.readCOO(str(path))
    got = bs.readCOO_gpu(str(path))
    assert got[2:] == want[2:]
    assert (got[0] == want[0]).all() and (got[1] == want[1]).all()
    Arow, Acol, N = got[0].astype(np.int32), got[1].astype(np.int32), got[3]
    check(bs, oracle, Acol, Arow, N, Acol, Arow, N, N)


@pytest.mark.parametrize("kernel", ["BSPGEMM_SORT_ASYNC", "BSPGEMM_SORT_SYNC"])
def test_sort_kernels_on_tiny_and_ragged_shapes(bs, oracle, monkeypatch, kernel):
    """Both sorting-network kernels (fused_sort.cuh) where the tile pipeline has nothing to chew on: fewer rows than one tile,
    fewer tiles than warps, a last tile that is cut short, empty A rows, rectangular B."""
    monkeypatch.setenv("BSPGEMM_FORCE_ELL", "1")
    monkeypatch.setenv("BSPGEMM_FORCE_SORT", "1")
    monkeypatch.setenv(kernel, "1")
    rng = np.random.default_rng(91)
    for An, Bn, Bm, da, db in ((1, 40, 1000, 16, 16), (3, 64, 1 << 20, 16, 16), (5, 300, 70000, 12, 14), (17, 1000, 1 << 22, 16, 16),
                               (149, 500, 5000, 7, 6), (2663, 4000, 1 << 21, 16, 16), (4099, 9000, 123457, 30, 30)):
        Arow, Acol = random_csr(rng, An, Bn, float(da), sort=False, dups=True)
        Brow, Bcol = random_csr(rng, Bn, Bm, float(db), sort=False, dups=False)
        if np.diff(Brow).max() > 32 or np.diff(Arow).max() > 32:
            keep_b = np.minimum(np.diff(Brow), 32)
            Bcol = np.concatenate([Bcol[Brow[i]:Brow[i] + keep_b[i]] for i in range(Bn)]).astype(np.int32)
            Brow = np.concatenate([[0], np.cumsum(keep_b)]).astype(np.int32)
            keep_a = np.minimum(np.diff(Arow), 32)
            Acol = np.concatenate([Acol[Arow[i]:Arow[i] + keep_a[i]] for i in range(An)]).astype(np.int32)
            Arow = np.concatenate([[0], np.cumsum(keep_a)]).astype(np.int32)
        want_col, want_row = oracle.spgemm(Acol, Arow, An, Bcol, Brow, Bm)
        for i64 in (False, True):
            got_col, got_row, st = dev_multiply(bs, bs.MODE_AUTO, Acol, Arow, An, Bcol, Brow, Bn, Bm, i64=i64)
            msg = _explain(got_col, got_row, want_col, want_row)
            assert not msg, f"An={An} Bn={Bn} Bm={Bm} [{kernel}] i64={i64} variant={st['variant']}: {msg}"


# ------------------------------------------------------------------------------------------------ prepared B / cached plan
def _dev_handle_product(bs, h, dAc, dAr, An, Annz, dBc, dBr, Bn, Bm, Bnnz, i64=False):
    import torch
    dCr = torch.full((An + 1,), -7, dtype=torch.int64 if i64 else torch.int32, device=dAr.device)
    ptr, nnz = h.multiply(dAc, dAr, An, Annz, dBc, dBr, Bn, Bm, Bnnz, dCr, crow_is_i64=i64)
    torch.cuda.synchronize()
    return bs.device_view(ptr, nnz, 0).cpu().numpy().copy(), dCr.cpu().numpy(), h.stats()


def test_prepared_b_gives_identical_results_and_replays_the_plan(bs, oracle):
    """bspgemm_dev_prepare_b: the ELL copy of B and the plan survive across products.  First product: probes, no re-layout;
    following products: launched from the cached plan (one kernel).  Every result is compared with the oracle; then A changes
    to rows LONGER than the plan's LA (detected on the device, redone with fresh probes), to ragged rows (host-side rule sends
    it the long way), and back."""
    import torch
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
    n = 30011
    brow, bcol = bs.gen_uniform(n, 8, 3)
    dBr, dBc = t(brow), t(bcol)
    h = bs.DeviceSpGEMM(0)
    h.prepare_b(dBc, dBr, n, n, len(bcol))
    rng = np.random.default_rng(5)

    def product(arow, acol, An, expect_cached, expect_variant=2):
        want_col, want_row = oracle.spgemm(acol, arow, An, bcol, brow, n)
        got_col, got_row, st = _dev_handle_product(bs, h, t(acol), t(arow), An, len(acol), dBc, dBr, n, n, len(bcol))
        msg = _explain(got_col, got_row, want_col, want_row)
        assert not msg, msg
        assert st["b_prepared"] == 1 and st["plan_cached"] == (1 if expect_cached else 0), st
        if expect_variant is not None:
            assert st["variant"] == expect_variant, st
        return st

    a8r, a8c = bs.gen_uniform(n, 8, 11)
    st = product(a8r, a8c, n, False)                 # first product with this B: probes run, the ELL copy is reused
    assert st["launches"] == 3                       # k_maxlen (A only), k_probe_span, the sort kernel: no k_build_ell in this product
    st = product(a8r, a8c, n, True)                  # replayed
    assert st["launches"] == 1
    a8r2, a8c2 = bs.gen_uniform(n - 7, 8, 12)        # another A of the same kind (fewer rows): still replayed
    product(a8r2, a8c2, n - 7, True)
    a16r, a16c = bs.gen_uniform(n, 16, 13)           # rows longer than the plan's LA = 8: flagged by the kernel, redone
    product(a16r, a16c, n, False)
    product(a16r, a16c, n, True)                     # ... and the new plan (LA = 16) is cached
    product(a8r, a8c, n, False)                      # half-length rows are not "regular" for the LA = 16 plan (host-side rule): fresh probes, LA = 8 again
    product(a8r, a8c, n, True)
    rr, rc = random_csr(rng, 5000, n, 3.0)           # ragged rows: not "regular" -> host-side rule declines the replay
    product(rr, rc, 5000, False, expect_variant=None)
    # an unrelated B through the same handle invalidates nothing silently: its product is correct, and so is the next prepared one
    b2r, b2c = bs.gen_uniform(n, 4, 21)
    want_col, want_row = oracle.spgemm(a8c, a8r, n, b2c, b2r, n)
    got_col, got_row, st = _dev_handle_product(bs, h, t(a8c), t(a8r), n, len(a8c), t(b2c), t(b2r), n, n, len(b2c))
    assert not _explain(got_col, got_row, want_col, want_row) and st["b_prepared"] == 0
    product(a8r, a8c, n, False, expect_variant=None)  # the ELL buffer was overwritten: rebuilt, result still exact
    h.forget_b()
    want_col, want_row = oracle.spgemm(a8c, a8r, n, bcol, brow, n)
    got_col, got_row, st = _dev_handle_product(bs, h, t(a8c), t(a8r), n, len(a8c), dBc, dBr, n, n, len(bcol))
    assert not _explain(got_col, got_row, want_col, want_row) and st["b_prepared"] == 0 and st["plan_cached"] == 0
    h.close()


def test_prepared_b_band_plan(bs, oracle):
    """Banded B: prepare_b keeps the run descriptors; the replayed product is the band kernel alone.  An A whose output rows are
    wider than the register bitmap fails on the device and is redone by the general kernels."""
    import torch
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
    n = 20000
    brow, bcol = bs.gen_banded(n, 32)
    dBr, dBc = t(brow), t(bcol)
    h = bs.DeviceSpGEMM(0)
    h.prepare_b(dBc, dBr, n, n, len(bcol))
    for k, cached in enumerate((False, True, True)):
        want_col, want_row = oracle.spgemm(bcol, brow, n, bcol, brow, n)
        got_col, got_row, st = _dev_handle_product(bs, h, dBc, dBr, n, len(bcol), dBc, dBr, n, n, len(bcol))
        assert not _explain(got_col, got_row, want_col, want_row)
        assert st["variant"] == 3 and st["plan_cached"] == (1 if cached else 0) and st["b_prepared"] == 1, st
    assert st["launches"] == 1
    wr, wc = bs.gen_uniform(n, 4, 9)                  # scattered A: output rows span far more than 128 columns
    want_col, want_row = oracle.spgemm(wc, wr, n, bcol, brow, n)
    got_col, got_row, st = _dev_handle_product(bs, h, t(wc), t(wr), n, len(wc), dBc, dBr, n, n, len(bcol))
    assert not _explain(got_col, got_row, want_col, want_row)
    assert st["plan_cached"] == 0 and st["variant"] != 3
    h.close()


def test_out_of_range_a_column_equal_to_bn_is_reported(bs):
    """An A entry equal to Bn (the internal "no row" sentinel of the ELL kernels) is out of range like any other."""
    import os
    n = 4096
    row, col = bs.gen_uniform(n, 8, 1)
    bad = col.copy(); bad[len(bad) // 2] = n
    for env in ({}, {"BSPGEMM_SORT_ASYNC": "1"}, {"BSPGEMM_NO_SORT": "1"}, {"BSPGEMM_NO_ELL": "1"}):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            with pytest.raises(bs.BSpGEMMError) as e:
                dev_multiply(bs, bs.MODE_AUTO, bad, row, n, col, row, n, n)
            assert e.value.status == bs.ERR_BADARG, env
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v


# ------------------------------------------------------------------------------------------------ BASELINE configs 4 and 5 at full size
def _check_blocks_vs_oracle(bs, oracle, row, col, n, C, Cr, blocks):
    """Bit-exact compare of contiguous row blocks of a device-resident result (C: int32 tensor, Cr: int64 row pointers)."""
    for r0, r1 in blocks:
        wc, wr = oracle.spgemm(col, row[r0:], r1 - r0, col, row, n)
        gr = Cr[r0:r1 + 1].cpu().numpy()
        assert (gr - gr[0] == wr).all(), f"row pointers differ in rows [{r0},{r1})"
        got = C[int(gr[0]):int(gr[-1])].cpu().numpy()
        assert len(got) == len(wc) and (got == wc).all(), f"columns differ in rows [{r0},{r1})"


def _chunked_contract(C, Cr, nnz, n_cols, chunk=1 << 28):
    """Result contract on a device-resident CSR too large for one-shot temporaries: columns in range, strictly ascending
    inside every row (a descent is allowed only at a row boundary)."""
    import torch
    assert int(Cr[0]) == 0 and int(Cr[-1]) == nnz and bool((Cr[1:] >= Cr[:-1]).all())
    starts = Cr[1:-1]
    starts = starts[(starts > 0) & (starts < nnz)]
    boundary = torch.zeros(0, dtype=torch.bool, device=C.device)
    for a in range(0, nnz, chunk):
        b = min(nnz, a + chunk)
        x = C[a:b]
        assert int(x.min()) >= 0 and int(x.max()) < n_cols
        hi = min(nnz, b + 1)
        d = C[a + 1:hi] > C[a:hi - 1]                        # d[p-a] : C[p+1] > C[p]
        s = starts[(starts > a) & (starts <= hi - 1)] - 1 - a   # descents allowed at p = start-1
        d[s] = True
        assert bool(d.all()), f"a row is not strictly ascending in entries [{a},{hi})"
        del d, x


def test_config5_full_size_closed_form_and_oracle_blocks(bs, oracle):
    """BASELINE config 5: banded n=2^24, d=32 (row i = columns i-16..i+15 clipped), C = A·A through bspgemm_dev_multiply (band
    kernel, 131072 tiles).  The WHOLE output is compared with the closed form (row i of C = columns i-32..i+30 clipped), and the
    closed form itself is pinned by the oracle on blocks at the start, in the middle and at the end."""
    import torch
    n = 1 << 24
    row, col = bs.gen_banded(n, 32)
    dev = torch.device("cuda:0")
    dAr, dAc = torch.from_numpy(row).to(dev), torch.from_numpy(col).to(dev)
    h = bs.DeviceSpGEMM(0)
    dCr = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    ptr, nnz = h.multiply(dAc, dAr, n, len(col), dAc, dAr, n, n, len(col), dCr)
    torch.cuda.synchronize()
    st = h.stats()
    assert st["variant"] == 3, st                                        # the band kernel ran
    i = torch.arange(n, device=dev, dtype=torch.int64)
    lo = torch.clamp(i - 32, min=0)
    hi = torch.clamp(i + 30, max=n - 1)
    want_row = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    want_row[1:] = torch.cumsum(hi - lo + 1, 0)
    Cr = dCr.to(torch.int64)
    assert torch.equal(Cr, want_row) and nnz == int(want_row[-1])
    C = bs.device_view(ptr, nnz, 0)
    for a in range(0, n, 1 << 22):                                       # closed-form columns, 2^22 rows at a time
        b = min(n, a + (1 << 22))
        lens = (hi - lo + 1)[a:b]
        base = torch.repeat_interleave(lo[a:b] - want_row[a:b], lens)    # column = lo[row] + (p - Crow[row])
        p0, p1 = int(want_row[a]), int(want_row[b])
        want = (base + torch.arange(p0, p1, device=dev, dtype=torch.int64)).to(torch.int32)
        assert torch.equal(C[p0:p1], want), f"columns differ from the closed form in rows [{a},{b})"
        del base, want
    _check_blocks_vs_oracle(bs, oracle, row, col, n, C, Cr, [(0, 20000), (n // 2 - 10000, n // 2 + 10000), (n - 20000, n)])
    assert st["ip"] == oracle.intermediate_products(col, row, n, row)
    h.close()


def test_config4_full_size_rmat_scale22_sampled_blocks(bs, oracle):
    """BASELINE config 4: R-MAT (.45,.22,.22,.11) scale 22, edge factor 16, C = A·A with 64-bit row pointers (nnz(C) = 1.15e10,
    46 GB of columns; staging arena of the big rows, windowed bitmaps, CTA-wide sorts, int64 chain values).  Contract checks on
    the whole output in chunks, the totals against the oracle's IP count, and bit-exact row blocks against the 64-bit oracle:
    the hub rows at the start, the middle, the end and scattered rows."""
    import gc
    import torch
    gc.collect()
    torch.cuda.empty_cache()                  # blocks cached by the earlier tests count as used in mem_get_info
    free, total = torch.cuda.mem_get_info(0)
    if free < 120 * (1 << 30):                # 46 GB of output + 48 GB of staging arena + inputs, lists and headroom
        pytest.skip(f"needs ~120 GB of free device memory, {free >> 30} GB are free")
    n = 1 << 22
    row, col = bs.gen_rmat(22, 16, 0.45, 0.22, 0.22, 1)
    dev = torch.device("cuda:0")
    dAr, dAc = torch.from_numpy(row).to(dev), torch.from_numpy(col).to(dev)
    h = bs.DeviceSpGEMM(0)
    dCr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    ptr, nnz = h.multiply(dAc, dAr, n, len(col), dAc, dAr, n, n, len(col), dCr, crow_is_i64=True)
    torch.cuda.synchronize()
    st = h.stats()
    assert nnz > (1 << 31) and st["rows_l"] > 0 and st["rows_m"] > 0, st     # every bin, and beyond the 32-bit ABI
    assert st["ip"] == oracle.intermediate_products(col, row, n, row) and nnz <= st["ip"]
    C = bs.device_view(ptr, nnz, 0)
    _chunked_contract(C, dCr, nnz, n)
    rng = np.random.default_rng(4)
    blocks = [(0, 600), (n // 2 - 4000, n // 2 + 4000), (n - 8000, n)]
    blocks += [(int(r), int(r) + 1) for r in rng.integers(0, n, 24)]
    lens = np.diff(row)
    blocks += [(int(r), int(r) + 1) for r in np.argsort(lens)[-3:]]          # the three longest rows of A
    _check_blocks_vs_oracle(bs, oracle, row, col, n, C, dCr, blocks)
    h.close()


# ------------------------------------------------------------------------------------------------ masked product, iterated products (N4)
def test_masked_product_matches_the_oracle(gpu_ctx, oracle, bs):
    """C = F .* (A·B) (final/SpGEMM_mpi_omp.c:232-288) through the host-pointer operator and the device operator: sorted masks,
    unsorted masks with repeats (canonicalised on the device), mask = A (the triangle pattern), empty masks, rectangular
    shapes, skewed rows; bit-exact against the oracle's restatement (itself pinned by the compiled reference)."""
    import torch
    rng = np.random.default_rng(31)
    cases = []
    n = 20011
    ar, ac = bs.gen_uniform(n, 8, 5)
    fr, fc = random_csr(rng, n, n, 200.0)
    cases.append(("uniform, sorted mask", ac, ar, n, ac, ar, n, n, fc, fr))
    fr2, fc2 = random_csr(rng, n, n, 60.0, sort=False, dups=True)
    cases.append(("uniform, unsorted mask with repeats", ac, ar, n, ac, ar, n, n, fc2, fr2))
    cases.append(("mask = A", ac, ar, n, ac, ar, n, n, ac, ar))
    cases.append(("empty mask", ac, ar, n, ac, ar, n, n, np.zeros(0, np.int32), np.zeros(n + 1, np.int32)))
    rr, rc = bs.gen_rmat(12, 16, 0.45, 0.22, 0.22, 7)
    m = 1 << 12
    cases.append(("rmat, mask = A", rc, rr, m, rc, rr, m, m, rc, rr))
    a2r, a2c = random_csr(rng, 300, 500, 7.0, sort=False, dups=True)
    b2r, b2c = random_csr(rng, 500, 900, 9.0, sort=False, dups=True)
    f2r, f2c = random_csr(rng, 300, 900, 90.0, sort=False, dups=True)
    cases.append(("rectangular", a2c, a2r, 300, b2c, b2r, 500, 900, f2c, f2r))
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
    h = bs.DeviceSpGEMM(0)
    for name, Acol, Arow, An, Bcol, Brow, Bn, Bm, Fcol, Frow in cases:
        want_col, want_row = oracle.spgemm_masked(Acol, Arow, An, Bcol, Brow, Bm, Fcol, Frow)
        got_col, got_row = bs.spgemm_csr_masked(Acol, Arow, An, Bcol, Brow, Bn, Bm, Fcol, Frow)
        msg = _explain(got_col, got_row, want_col, want_row)
        assert not msg, (name, msg)
        for i64 in (False, True):
            dCr = torch.full((An + 1,), -3, dtype=torch.int64 if i64 else torch.int32, device=dev)
            ptr, nnz = h.multiply_masked(t(Acol), t(Arow), An, len(Acol), t(Bcol), t(Brow), Bn, Bm, len(Bcol), t(Fcol), t(Frow), len(Fcol), dCr, crow_is_i64=i64)
            torch.cuda.synchronize()
            msg = _explain(bs.device_view(ptr, nnz, 0).cpu().numpy(), dCr.cpu().numpy(), want_col, want_row)
            assert not msg, (name, i64, msg)
    h.close()


def test_iterated_products_stay_on_the_device(bs, oracle):
    """Powers of a graph (the motivating application, report p.1: paths of length 2, 3, ...): A^2, then A^3 = A^2·A with the
    prepared B = A, every factor device-resident between the products (the handle's arena is copied to a tensor of the caller
    before the next product reuses it); then the masked closure step A .* A^3.  Compared with the oracle's chain."""
    import torch
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
    n = 6007
    ar, ac = bs.gen_uniform(n, 3, 17)
    dAr, dAc = t(ar), t(ac)
    h = bs.DeviceSpGEMM(0)
    h.prepare_b(dAc, dAr, n, n, len(ac))
    cur_r, cur_c, d_r, d_c = ar, ac, dAr, dAc
    for power in (2, 3, 4):
        want_c, want_r = oracle.spgemm(cur_c, cur_r, n, ac, ar, n)
        dCr = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        ptr, nnz = h.multiply(d_c, d_r, n, int(d_c.numel()), dAc, dAr, n, n, len(ac), dCr)
        torch.cuda.synchronize()
        d_c = bs.device_view(ptr, nnz, 0).clone()            # the next product overwrites the arena
        d_r = dCr
        assert h.stats()["b_prepared"] == 1
        msg = _explain(d_c.cpu().numpy(), d_r.cpu().numpy(), want_c, want_r)
        assert not msg, (power, msg)
        cur_r, cur_c = want_r.astype(np.int32), want_c
    want_c, want_r = oracle.spgemm_masked(cur_c, cur_r, n, ac, ar, n, ac, ar)          # A .* (A^4 · A)
    dCr = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    ptr, nnz = h.multiply_masked(d_c, d_r, n, int(d_c.numel()), dAc, dAr, n, n, len(ac), dAc, dAr, len(ac), dCr)
    torch.cuda.synchronize()
    msg = _explain(bs.device_view(ptr, nnz, 0).cpu().numpy(), dCr.cpu().numpy(), want_c, want_r)
    assert not msg, msg
    h.close()


def test_sprand_cheap_rows_two_unordered_passes(bs, oracle):
    """The report's matrices (Matlab sprand(n,n,d/n)>0, Poisson(d) rows; /root/reference/Matlab/write_spm.m:5): every row is small and
    cheap, so the product runs as count -> device scan -> fill (no ordered one-pass kernel), with the warp bin sized from the
    measured maximum of the rows' products instead of the bound max_len(A) x max_len(B)."""
    n, d = 150_000, 5.0
    row, col = bs.gen_sprand(n, d, 3)
    want_col, want_row = oracle.spgemm(col, row, n, col, row, n)
    got_col, got_row, st = dev_multiply(bs, bs.MODE_FUSED, col, row, n, col, row, n, n)
    msg = _explain(got_col, got_row, want_col, want_row)
    assert not msg, msg
    assert st["rows_m"] == 0 and st["rows_l"] == 0 and st["cap_s"] <= 256, st
