"""tools/time_coo2csc.py — GPU coo2csc (bspgemm_coo2csc_dev, device-resident) vs the host coo2csc at BASELINE sizes."""
import importlib, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
bs = importlib.import_module("binary-spgemm_b200")
L = bs.lib()
for name, n, d in (("cfg2", 1 << 20, 8), ("cfg3", 1 << 22, 16)):
    rng = np.random.default_rng(1)
    nnz = n * d
    I = rng.integers(0, n, nnz, dtype=np.uint32); J = rng.integers(0, n, nnz, dtype=np.uint32)
    t0 = time.perf_counter(); hr, hc = bs.coo2csc(I, J, n, 0); t_host = time.perf_counter() - t0
    dev = torch.device("cuda:0")
    dI = torch.from_numpy(I.view(np.int32)).to(dev); dJ = torch.from_numpy(J.view(np.int32)).to(dev)
    dr = torch.empty(nnz, dtype=torch.int32, device=dev); dc = torch.empty(n + 1, dtype=torch.int32, device=dev)
    ms = []
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        st = L.bspgemm_coo2csc_dev(None, dr.data_ptr(), dc.data_ptr(), dI.data_ptr(), dJ.data_ptr(), nnz, n, 0)
        torch.cuda.synchronize(); ms.append((time.perf_counter() - t0) * 1e3)
        assert st == 0, L.bspgemm_last_error()
    ok = bool((dr.cpu().numpy().view(np.uint32) == hr).all() and (dc.cpu().numpy().view(np.uint32) == hc).all())
    print(f"{name}: n={n} nnz={nnz}  host coo2csc {t_host*1e3:.1f} ms   GPU (device-resident, incl. its cudaMalloc/cudaFree) "
          f"{min(ms[1:]):.2f} ms   identical={ok}")
