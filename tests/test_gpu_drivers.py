"""The drop-in executables and the in-process multi-task path, on the GPU (b).

* host/SpGEMM_gpu and host/SpGEMM_gpu_validity are executed on the reference's fixture and their stdout is pinned byte for
  byte: exactly the 11-field line of final/SpGEMM_mpi_omp.c:336 / exactly the string of final/SpGEMM_mpi_omp_validity.c:340.
* oracle/_ref/SpGEMM_mpi_omp*_bspgemm are the REFERENCE'S OWN drivers (its main, its readCOO, its timing loop, its serial
  SpGEMM_bigslice check) with the SpGEMM_mpi call sites switched to libbspgemm.so: bspgemm_SpGEMM_mpi must
  be an undefined symbol of the executable, and the reference's own validity check must accept the GPU result.
* bspgemm_csr with several tasks (row-block shards, displacements, gather with the offset applied on the device) is compared
  with the oracle: on several GPUs when the box has them, and always with several shards sharing GPU 0."""
import os
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from helpers import random_csr

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parents[1]
HOSTDIR = ROOT / "binary-spgemm_b200" / "host"
REFDIR = ROOT / "oracle" / "_ref"
FLT = r"\d+\.\d{6}"


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def fixture_mtx(bs, fixture_npz, tmp_path_factory):
    """The reference's validity_test.mtx: the copy next to the compiled reference, else rewritten from the golden arrays."""
    p = REFDIR / "validity_test.mtx"
    if p.exists():
        return str(p)
    q = tmp_path_factory.mktemp("mtx") / "validity_test.mtx"
    bs.write_mtx(str(q), fixture_npz["Arow"], fixture_npz["Acol"])
    return str(q)


def _run(cmd, **env):
    e = dict(os.environ)
    e.pop("BSPGEMM_GPUS", None); e.pop("BSPGEMM_DEVICES", None)
    e.update({k: str(v) for k, v in env.items()})
    return subprocess.run([str(c) for c in cmd], capture_output=True, text=True, env=e, timeout=300)


def _csv_re(tasks, threads, block, path):
    return re.compile(rf"^{tasks},{threads},{tasks * threads},{block},{re.escape(path)},50000,25000,12502,{FLT},{FLT},{FLT}\n$")


@pytest.mark.parametrize("tasks", [1, 3])
def test_perf_driver_stdout_is_exactly_the_reference_line(bs, fixture_mtx, tasks):
    """final/SpGEMM_mpi_omp.c:336 — one line, 11 fields, %lf = 6 decimals; nothing else on stdout (NCCL's banner, the
    second metrics line and every diagnostic go to stderr).  tasks = 3: three row-block shards sharing GPU 0."""
    env = {"BSPGEMM_DEVICES": ",".join(["0"] * tasks)} if tasks > 1 else {}
    r = _run([HOSTDIR / "SpGEMM_gpu", fixture_mtx, 6250, 2, 3], **env)
    assert r.returncode == 0, r.stderr
    assert _csv_re(tasks, 2, 6250, fixture_mtx).match(r.stdout), repr(r.stdout)
    assert "ip=12502" in r.stderr


def test_perf_driver_gpu_reader_hook(bs, fixture_mtx):
    r = _run([HOSTDIR / "SpGEMM_gpu", fixture_mtx, 6250, 8, 1], BSPGEMM_GPU_COO2CSC=1)
    assert r.returncode == 0, r.stderr
    assert _csv_re(1, 8, 6250, fixture_mtx).match(r.stdout), repr(r.stdout)


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_perf_driver_two_gpus(bs, fixture_mtx):
    r = _run([HOSTDIR / "SpGEMM_gpu", fixture_mtx, 6250, 2, 2], BSPGEMM_GPUS=2)
    assert r.returncode == 0, r.stderr
    assert _csv_re(2, 2, 6250, fixture_mtx).match(r.stdout), repr(r.stdout)


def test_usage_errors_like_the_reference(bs):
    r = _run([HOSTDIR / "SpGEMM_gpu", "x.mtx"])                      # argc != 5 -> usage + exit(1) (:357-360)
    assert r.returncode == 1 and r.stdout.startswith("usage:")
    r = _run([HOSTDIR / "SpGEMM_gpu", "/nonexistent/x.mtx", 1, 1, 1])   # unreadable file -> exit(1) (final/utils.c:54-61)
    assert r.returncode == 1


@pytest.mark.parametrize("devices", ["0", "0,0,0,0"])
def test_validity_driver_make_test(bs, fixture_mtx, devices):
    """`make test` (final/Makefile:11-12: 4 tasks x 2 threads x 6250): the reference's string, exit status 0."""
    r = _run([HOSTDIR / "SpGEMM_gpu_validity", fixture_mtx, 6250, 2], BSPGEMM_DEVICES=devices)
    assert r.returncode == 0, r.stderr
    assert r.stdout == "Results of serial and multricore are the same!\n", repr(r.stdout)


def test_make_test_target(bs):
    if not (REFDIR / "validity_test.mtx").exists():
        pytest.skip("fixture file not present")
    r = subprocess.run(["make", "-s", "-C", str(ROOT / "binary-spgemm_b200"), "test"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout.endswith("Results of serial and multricore are the same!\n"), repr(r.stdout)


# ------------------------------------------------------------------------------------------------ the reference's own main on the GPU
def _ref_gpu_binary(name):
    p = REFDIR / name
    if not p.exists():
        pytest.skip(f"{p} not built (needs /root/reference at build time; the prebuilt file travels to the GPU box)")
    return p


def test_reference_driver_calls_the_library_not_its_own_cpu_code(bs):
    """nm: bspgemm_SpGEMM_mpi is UNDEFINED in the reference's executable (resolved from libbspgemm.so at load time); the
    reference's CPU SpGEMM_mpi is still defined there but no longer called from test_mpi."""
    for name in ("SpGEMM_mpi_omp_bspgemm", "SpGEMM_mpi_omp_validity_bspgemm"):
        p = _ref_gpu_binary(name)
        und = subprocess.run(["nm", "-u", str(p)], capture_output=True, text=True).stdout
        assert re.search(r"^\s+U bspgemm_SpGEMM_mpi$", und, re.M), und
        defined = subprocess.run(["nm", "--defined-only", str(p)], capture_output=True, text=True).stdout
        assert "bspgemm_" not in defined
        ldd = subprocess.run(["ldd", str(p)], capture_output=True, text=True).stdout
        assert "libbspgemm.so" in ldd


@pytest.mark.parametrize("devices", ["0", "0,0"])
def test_reference_main_drives_the_gpu_kernels(bs, fixture_mtx, devices):
    """The reference's perf driver (its main / readCOO / timing loop / printf, unmodified but for the call-site macro) on the
    GPU library: its own CSV line with the reference's nnz(C) = 12502.  (tasks = MPI ranks of the stub = 1.)"""
    p = _ref_gpu_binary("SpGEMM_mpi_omp_bspgemm")
    r = _run([p, fixture_mtx, 6250, 2, 2], BSPGEMM_DEVICES=devices)
    assert r.returncode == 0, r.stderr
    assert _csv_re(1, 2, 6250, fixture_mtx).match(r.stdout), repr(r.stdout)


@pytest.mark.parametrize("devices", ["0", "0,0,0"])
def test_reference_validity_check_accepts_the_gpu_result(bs, fixture_mtx, devices):
    """The reference's validity driver: SpGEMM_mpi -> GPU, then the REFERENCE'S serial CPU SpGEMM_bigslice over all rows and its
    SpGEMM_valid element-wise compare (final/SpGEMM_mpi_omp_validity.c:331-343) — the reference itself judges the GPU output."""
    p = _ref_gpu_binary("SpGEMM_mpi_omp_validity_bspgemm")
    r = _run([p, fixture_mtx, 6250, 2], BSPGEMM_DEVICES=devices)
    assert r.returncode == 0, r.stderr
    assert r.stdout == "Results of serial and multricore are the same!\n", repr(r.stdout)


# ------------------------------------------------------------------------------------------------ bspgemm_csr with several tasks vs the oracle
def _cases(bs):
    rng = np.random.default_rng(77)
    row, col = bs.gen_uniform(40001, 8, 5)                       # An not divisible by the task count; ELL sort kernel
    yield "uniform", col, row, 40001, col, row, 40001, 40001
    rrow, rcol = bs.gen_rmat(13, 16, 0.45, 0.22, 0.22, 3)        # skewed rows: every bin, equal-row shards of unequal work
    yield "rmat13", rcol, rrow, 1 << 13, rcol, rrow, 1 << 13, 1 << 13
    ar, ac = random_csr(rng, 5, 300, 6, sort=False, dups=True)   # fewer rows than tasks: empty shards join the broadcast
    br, bc = random_csr(rng, 300, 77, 9, sort=False, dups=True)
    yield "An<tasks", ac, ar, 5, bc, br, 300, 77
    yield "An=1", ac[ar[0]:ar[1]], ar[:2], 1, bc, br, 300, 77
    yield "An=0", np.zeros(0, np.int32), np.zeros(1, np.int32), 0, bc, br, 300, 77
    yield "empty rows", np.zeros(0, np.int32), np.zeros(12, np.int32), 11, bc, br, 300, 77


def _multi_task_check(bs, oracle, devices):
    bs.finalize()
    try:
        bs.init(devices=devices)
        assert bs.num_gpus() == len(devices)
        for name, Acol, Arow, An, Bcol, Brow, Bn, Bm in _cases(bs):
            want_col, want_row = oracle.spgemm(Acol, Arow, An, Bcol, Brow, Bm)
            for i64 in (False, True):
                got_col, got_row = bs.spgemm_csr(Acol, Arow, An, Bcol, Brow, Bn, Bm, i64=i64)
                assert (np.asarray(got_row, np.int64) == want_row).all(), (name, devices, i64)
                assert (got_col == want_col).all(), (name, devices, i64)
            buf = np.full(len(want_col) + 3, -1, np.int32)
            nnz, crow = bs.spgemm_csr_into(Acol, Arow, An, Bcol, Brow, Bn, Bm, buf)
            assert nnz == len(want_col) and (buf[:nnz] == want_col).all() and (crow == want_row).all(), name
    finally:
        bs.finalize()
        bs.init(1)


@pytest.mark.parametrize("tasks", [2, 3, 8])
def test_sharded_host_operator_tasks_sharing_one_gpu(gpu_ctx, oracle, tasks):
    """Row-block split (final/SpGEMM_mpi_omp.c:165-171), per-task products, displacement scan (:189-196) and the gather with the
    row-pointer offset applied on the device (k_offset_rowptr, replaces :211-223) — several tasks on GPU 0, no communicator."""
    _multi_task_check(gpu_ctx, oracle, [0] * tasks)


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_sharded_host_operator_over_nccl(gpu_ctx, oracle):
    """The same with one task per GPU: B uploaded once and replicated by ncclBroadcast over NVLink; includes An < #GPUs."""
    _multi_task_check(gpu_ctx, oracle, list(range(min(_ngpu(), 8))))


# ------------------------------------------------------------------------------------------------ distributed consumer (N3)
def _sharded_consumer_check(bs, oracle, devices, tmp_path):
    import torch
    rng = np.random.default_rng(123)
    n = 30007
    row, col = bs.gen_uniform(n, 8, 9)
    want_col, want_row = oracle.spgemm(col, row, n, col, row, n)
    bs.finalize()
    try:
        bs.init(devices=devices)
        for i64 in (False, True):
            R = bs.ShardedResult(col, row, n, col, row, n, n, i64=i64)
            assert R.nshards == len(devices) and R.nnz == len(want_col)
            rdt = torch.int64 if i64 else torch.int32
            # (a) the shards where they were computed: slice-relative row pointers, displacements = the exclusive scan of the shard sizes
            for q in range(R.nshards):
                s = R.shard(q)
                assert s["device"] == devices[q] and s["row0"] == (n * q) // len(devices) and s["disp"] == want_row[s["row0"]]
                with torch.cuda.device(s["device"]):
                    cr = bs.device_view(s["dCrow"], (s["rows"] + 1) * (2 if i64 else 1), s["device"]).cpu().numpy().view(np.int64 if i64 else np.int32)
                    cc = bs.device_view(s["dCcol"], s["nnz"], s["device"]).cpu().numpy()
                assert (cr == want_row[s["row0"]:s["row0"] + s["rows"] + 1] - want_row[s["row0"]]).all()
                assert (cc == want_col[s["disp"]:s["disp"] + s["nnz"]]).all()
            # (b) all-gather on the devices: every task holds the whole CSR, row pointers already offset
            cols, rows = R.allgather()
            for q in range(R.nshards):
                with torch.cuda.device(devices[q]):
                    fr = bs.device_view(rows[q], (n + 1) * (2 if i64 else 1), devices[q]).cpu().numpy().view(np.int64 if i64 else np.int32)
                    fc = bs.device_view(cols[q], R.nnz, devices[q]).cpu().numpy()
                assert (fr == want_row).all() and (fc == want_col).all(), (q, i64)
            # (c) one file per shard
            R.write(str(tmp_path / f"c{int(i64)}"), 0)
            R.write(str(tmp_path / f"c{int(i64)}"), 1)
            for q in range(R.nshards):
                raw = np.fromfile(tmp_path / f"c{int(i64)}.shard{q}.bin", dtype=np.uint8)
                hdr = raw[:64].view(np.int64)
                assert hdr[0] == 0x3152534347505342 and hdr[1] == n and hdr[2] == n
                r0, nr, nz, dp = (int(x) for x in hdr[3:7])
                pr = raw[64:64 + 8 * (nr + 1)].view(np.int64)
                pc = raw[64 + 8 * (nr + 1):].view(np.int32)
                assert dp == want_row[r0] and (pr == want_row[r0:r0 + nr + 1] - want_row[r0]).all() and len(pc) == nz
                assert (pc == want_col[dp:dp + nz]).all()
                mr, mc, M, N, nnz = bs.readCOO(str(tmp_path / f"c{int(i64)}.shard{q}.mtx"))      # the shard alone, global dimensions
                assert M == n and N == n and nnz == nz
                assert (np.diff(mr.astype(np.int64))[r0:r0 + nr] == np.diff(want_row[r0:r0 + nr + 1])).all() and mr[r0] == 0
                assert (mc.astype(np.int32) == pc).all()
            R.free()
    finally:
        bs.finalize()
        bs.init(1)


def test_distributed_consumer_tasks_sharing_one_gpu(gpu_ctx, oracle, tmp_path):
    """Shards left on the device, all-gather (device-to-device copies: no communicator), one file per shard."""
    _sharded_consumer_check(gpu_ctx, oracle, [0, 0, 0], tmp_path)


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_distributed_consumer_over_nccl(gpu_ctx, oracle, tmp_path):
    """The same with one task per GPU: the all-gather is one grouped ncclBroadcast per shard over NVLink."""
    _sharded_consumer_check(gpu_ctx, oracle, list(range(min(_ngpu(), 4))), tmp_path)
