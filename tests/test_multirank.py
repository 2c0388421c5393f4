"""The N>1 path of bench.py on CPU: two gloo ranks shard A into contiguous row blocks (the reference's rank
split, final/SpGEMM_mpi_omp.c:165-171), B is broadcast from rank 0, shards are gathered for validation with
the displacement fix-up (:189-223).  The per-shard product is the oracle here (no GPU) — what is under test
is the host-side sharding / broadcast / gather logic the GPU ranks use unchanged."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import importlib
    import torch
    import torch.distributed as dist
    import bench
    from oracle.oracle import Oracle
    bs = importlib.import_module("binary-spgemm_b200")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 5000 + 3                                  # not divisible by the world size on purpose
    if rank == 0:
        row, col = bs.gen_uniform(n, 8, 5)
    else:
        row = col = None
    row_t, col_t = bench.broadcast_csr(row, col, n, src=0, device="cpu")
    r0, r1 = bench.shard_bounds(n, rank, world)
    O = Oracle()
    Arow = row_t.numpy()[r0:r1 + 1]
    Ccol, Crow = O.spgemm(col_t.numpy(), Arow, r1 - r0, col_t.numpy(), row_t.numpy(), n)
    full = bench.gather_shards(torch.from_numpy(Ccol), torch.from_numpy(Crow), n, rank, world, device="cpu")
    if rank == 0:
        wc, wr = O.spgemm(col, row, n, col, row, n)
        gc, gr = full
        ok = bool((gr.numpy() == wr).all() and (gc.numpy() == wc).all())
        Path(out).write_text("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_broadcast_gather(tmp_path):
    import torch.multiprocessing as mp
    out = tmp_path / "res.txt"
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"


def test_shard_bounds_cover_all_rows():
    sys.path.insert(0, str(ROOT))
    import bench
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            b = [bench.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1
