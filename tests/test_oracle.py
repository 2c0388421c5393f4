"""Pins the oracle (oracle/spgemm_oracle.c) against the reference's golden vectors and, where the compiled
reference is present (oracle/_ref), against the reference itself.  CPU only."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import crc, gen_case, random_csr

# crc32 of the little-endian int32 arrays, SURVEY.md §4 "Golden values for config 1"
FIXTURE_CRC = dict(Arow="0327ec62", Acol="3d773ff2", Crow="62b292d7", Ccol="77f3757b")


def test_fixture_golden_crc(fixture_npz):
    for k, v in FIXTURE_CRC.items():
        assert crc(fixture_npz[k]) == v
    assert len(fixture_npz["Ccol"]) == 12502 and fixture_npz["Crow"][-1] == 12502


def test_oracle_coo2csc_and_spgemm_on_fixture(oracle, fixture_npz):
    f = fixture_npz
    M = int(f["M"])
    # readCOO hands (indices=I, keys=J) to coo2csc (final/utils.c:77)
    idx, ptr = oracle.coo2csc(f["I"], f["J"], M)
    assert (ptr == f["Arow"]).all() and (idx == f["Acol"]).all()
    Ccol, Crow = oracle.spgemm(f["Acol"], f["Arow"], M, f["Acol"], f["Arow"], M)
    assert (Crow == f["Crow"]).all() and (Ccol == f["Ccol"]).all()
    assert oracle.intermediate_products(f["Acol"], f["Arow"], M, f["Arow"]) == 12502   # no duplicate is ever produced


def test_oracle_kats(oracle, kats):
    for k in kats:
        An = len(k["Arow"]) - 1
        Ccol, Crow = oracle.spgemm(k["Acol"], k["Arow"], An, k["Bcol"], k["Brow"], k["Bm"], nslices=2)
        assert Crow.tolist() == k["Crow"], k["name"]
        assert Ccol.tolist() == k["Ccol"], k["name"]


def test_kats_match_survey_expectations(kats):
    """The scipy-derived expectations written down in SURVEY.md §4 agree with the compiled reference."""
    by = {k["name"]: k for k in kats}
    assert by["9x9_AA"]["Crow"] == [0, 4, 6, 7, 8, 11, 11, 13, 18, 19]
    assert by["9x9_AA"]["Ccol"] == [1, 2, 4, 6, 0, 2, 2, 3, 1, 3, 4, 1, 7, 2, 5, 6, 7, 8, 3]
    assert by["8x8_AA"]["Crow"] == [0, 1, 5, 5, 8, 8, 9, 12, 14]
    assert by["8x8_AA"]["Ccol"] == [0, 0, 2, 5, 6, 0, 5, 7, 0, 0, 1, 5, 0, 3]
    assert by["4x4_AB"]["Crow"] == [0, 3, 7, 9, 10] and by["4x4_AB"]["Ccol"] == [0, 2, 3, 0, 1, 2, 3, 0, 1, 0]
    assert by["4x4_AA"]["Crow"] == [0, 3, 6, 10, 12] and by["4x4_AA"]["Ccol"] == [1, 2, 3, 0, 1, 2, 0, 1, 2, 3, 2, 3]


def test_oracle_seeded_cases(oracle, bs, seeded_cases):
    for c in seeded_cases:
        row, col = gen_case(bs, c)
        assert [crc(row), crc(col)] == c["in_crc"], c["name"]        # generator is deterministic
        n = c["n"]
        Ccol, Crow = oracle.spgemm(col, row, n, col, row, n)
        assert Crow[-1] == c["nnzC"] and crc(Crow) == c["Crow_crc"] and crc(Ccol) == c["Ccol_crc"], c["name"]


def test_oracle_vs_compiled_reference_random(oracle, ref):
    rng = np.random.default_rng(5)
    for trial in range(12):
        n, k, m = (int(x) for x in rng.integers(1, 400, 3))
        Arow, Acol = random_csr(rng, n, k, rng.uniform(0, 6), sort=trial % 2 == 0, dups=trial % 3 == 0)
        Brow, Bcol = random_csr(rng, k, m, rng.uniform(0, 6), sort=trial % 2 == 1, dups=trial % 3 == 1)
        Rc, Rr = ref.bigslice(Acol, Arow, n, Bcol, Brow, m)
        Oc, Or = oracle.spgemm(Acol, Arow, n, Bcol, Brow, m, nslices=3)
        assert (Or == Rr).all() and (Oc == Rc).all()
        assert oracle.intermediate_products(Acol, Arow, n, Brow) == int(sum(Brow[j + 1] - Brow[j] for j in Acol))


def test_oracle_vs_reference_hybrid_path(oracle, ref, bs):
    """SpGEMM_mpi -> SpGEMM_omp with tBlock slices (what `make test` runs) equals the oracle."""
    row, col = bs.gen_uniform(8192, 8, 3)
    ref.set_threads(4)
    Hc, Hr = ref.mpi(col, row, 8192, col, row, 8192, 1024)
    Oc, Or = oracle.spgemm(col, row, 8192, col, row, 8192)
    assert (Or == Hr).all() and (Oc == Hc).all()


def test_oracle_vs_scipy(oracle):
    rng = np.random.default_rng(11)
    for _ in range(5):
        n, k, m = (int(x) for x in rng.integers(50, 600, 3))
        Arow, Acol = random_csr(rng, n, k, 5)
        Brow, Bcol = random_csr(rng, k, m, 7)
        A = sp.csr_matrix((np.ones(len(Acol), np.int64), Acol, Arow), shape=(n, k))
        B = sp.csr_matrix((np.ones(len(Bcol), np.int64), Bcol, Brow), shape=(k, m))
        C = A @ B
        C.sort_indices()
        Oc, Or = oracle.spgemm(Acol, Arow, n, Bcol, Brow, m)
        assert (C.indptr == Or).all() and (C.indices == Oc).all()


def test_oracle_readCOO_matches_reference(oracle, ref, tmp_path):
    import os
    fx = os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "validity_test.mtx")
    if not os.path.exists(fx):
        pytest.skip("fixture copy not present")
    a, b = oracle.readCOO(fx), ref.readCOO(fx)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2:] == b[2:]


def test_reference_runs_with_several_tasks_through_the_shm_mpi_shim():
    """oracle/mpi_shm/mpi.h: the unmodified reference drivers with 4 forked tasks — the reference's own
    `make test` command line (mpirun -n 4 … validity_test.mtx 6250 2, final/Makefile:11-12) — agree with their serial run, and
    the performance driver's CSV line carries tasks=4 and the fixture's sizes (50000, 25000 -> 12502)."""
    import os
    import subprocess
    from pathlib import Path
    ref = Path(__file__).resolve().parents[1] / "oracle" / "_ref"
    if not (ref / "SpGEMM_mpi_omp_validity_shm").exists():
        pytest.skip("oracle/_ref/*_shm not built (make -C oracle ref)")
    env = dict(os.environ, MPI_SHIM_TASKS="4", OMP_NUM_THREADS="2")
    out = subprocess.run([str(ref / "SpGEMM_mpi_omp_validity_shm"), "validity_test.mtx", "6250", "2"], cwd=ref, env=env,
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "Results of serial and multricore are the same!" in out.stdout, out.stdout + out.stderr
    out = subprocess.run([str(ref / "SpGEMM_mpi_omp_shm"), "validity_test.mtx", "6250", "2", "2"], cwd=ref, env=env,
                         capture_output=True, text=True, timeout=120)
    f = out.stdout.strip().split(",")
    assert out.returncode == 0 and f[:4] == ["4", "2", "8", "6250"] and f[5:8] == ["50000", "25000", "12502"], out.stdout + out.stderr


def _masked_cases(rng):
    n = 700
    ar, ac = random_csr(rng, n, n, 6.0)
    br, bc = random_csr(rng, n, n, 5.0, sort=False, dups=True)
    fr, fc = random_csr(rng, n, n, 40.0)                                # sorted, distinct mask
    yield "square sorted mask", ac, ar, n, bc, br, n, fc, fr
    fr2, fc2 = random_csr(rng, n, n, 25.0, sort=False, dups=True)       # the reference's mask is a flag array: any order, repeats
    yield "square unsorted mask with repeats", ac, ar, n, bc, br, n, fc2, fr2
    yield "mask = A (triangle pattern)", ac, ar, n, ac, ar, n, ac, ar
    fr3 = np.zeros(n + 1, np.int32)
    yield "empty mask", ac, ar, n, bc, br, n, np.zeros(0, np.int32), fr3


def test_oracle_masked_matches_reference_and_scipy(oracle, ref):
    """SpGEMM_masked (final/SpGEMM_mpi_omp.c:232-288) is defined by the reference but never called by its drivers; the
    compiled reference function pins the restatement, scipy's (A@B).multiply(F) pins both."""
    rng = np.random.default_rng(2024)
    for name, ac, ar, An, bc, br, Bm, fc, fr in _masked_cases(rng):
        got_col, got_row = oracle.spgemm_masked(ac, ar, An, bc, br, Bm, fc, fr)
        ref_col, ref_row = ref.masked(ac, ar, An, bc, br, Bm, fc, fr)
        assert (got_row == ref_row).all() and (got_col == ref_col).all(), name
        A = sp.csr_matrix((np.ones(len(ac), np.int64), ac, ar), shape=(An, len(br) - 1))
        B = sp.csr_matrix((np.ones(len(bc), np.int64), bc, br), shape=(len(br) - 1, Bm))
        F = sp.csr_matrix((np.ones(len(fc), np.int64), fc, fr), shape=(An, Bm))
        A.sum_duplicates(); B.sum_duplicates(); F.sum_duplicates()
        P = (A @ B).multiply(F).tocsr()
        P.eliminate_zeros(); P.sort_indices()
        assert (P.indptr == got_row).all() and (P.indices == got_col).all(), name


def test_oracle_masked_rectangular(oracle):
    """Rectangular A (30 x 50) · B (50 x 90) with a 30 x 90 mask: beyond the reference (its flag array has An entries, :239)."""
    rng = np.random.default_rng(9)
    ar, ac = random_csr(rng, 30, 50, 7.0)
    br, bc = random_csr(rng, 50, 90, 9.0, sort=False)
    fr, fc = random_csr(rng, 30, 90, 30.0, sort=False, dups=True)
    got_col, got_row = oracle.spgemm_masked(ac, ar, 30, bc, br, 90, fc, fr)
    full_col, full_row = oracle.spgemm(ac, ar, 30, bc, br, 90)
    for i in range(30):
        want = sorted(set(full_col[full_row[i]:full_row[i + 1]]) & set(fc[fr[i]:fr[i + 1]]))
        assert got_col[got_row[i]:got_row[i + 1]].tolist() == want
