#!/usr/bin/env python3
"""tools/shard_sweep.py — kernel time of the config-3 product as a function of the shard size on ONE GPU (first n/k rows of A, full B,
B prepared): separates the kernel's fixed cost from the per-tile cost, i.e. what an N-GPU run sees per rank."""
import importlib, sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
bs = importlib.import_module("binary-spgemm_b200")
n, d = 1 << 22, 16
row, col = bs.gen_uniform(n, d, 1)
dev = torch.device("cuda:0")
d_row, d_col = torch.from_numpy(row).to(dev), torch.from_numpy(col).to(dev)
h = bs.DeviceSpGEMM(0)
d_crow = torch.zeros(n + 1, dtype=torch.int32, device=dev)
nnzA = len(col)
out = []
for k in (1, 2, 4, 8, 16, 32, 64):
    rows = n // k
    shard_nnz = int(row[rows])
    call = h.bound_multiply(d_col, d_row, rows, shard_nnz, d_col, d_row, n, n, nnzA, d_crow)
    call(); h.prepare_b(d_col, d_row, n, n, nnzA)
    for _ in range(5): call()
    ms, tot = [], []
    for _ in range(20):
        call(); s = h.stats(); ms.append(s["ms_main"]); tot.append(s["ms_total"])
    h.forget_b()
    out.append((k, rows, float(np.median(ms)), float(np.median(tot))))
    print("1/%-3d rows %8d kernel %.4f ms  (x%d = %.3f)  total %.4f" % (k, rows, out[-1][2], k, out[-1][2] * k, out[-1][3]), flush=True)
ks = np.array([1.0 / o[0] for o in out]); t = np.array([o[2] for o in out])
A = np.vstack([np.ones_like(ks), ks]).T
c, w = np.linalg.lstsq(A, t, rcond=None)[0]
print("fit: kernel_ms = %.4f + %.4f * (shard / n)" % (c, w))
