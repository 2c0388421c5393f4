#!/usr/bin/env python3
"""tools/shard_one.py <k> — a few config-3 products on the first n/k rows of A (B prepared), for ncu."""
import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
bs = importlib.import_module("binary-spgemm_b200")
k = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n, d = 1 << 22, 16
row, col = bs.gen_uniform(n, d, 1)
dev = torch.device("cuda:0")
d_row, d_col = torch.from_numpy(row).to(dev), torch.from_numpy(col).to(dev)
h = bs.DeviceSpGEMM(0)
d_crow = torch.zeros(n + 1, dtype=torch.int32, device=dev)
rows = n // k
call = h.bound_multiply(d_col, d_row, rows, int(row[rows]), d_col, d_row, n, n, len(col), d_crow)
call(); h.prepare_b(d_col, d_row, n, n, len(col))
for _ in range(6): call()
print("kernel ms", h.stats()["ms_main"])
