"""tools/ref_tasks_sweep.py [workload [times [PxT,PxT,...]]] — the UNMODIFIED reference (oracle/_ref/SpGEMM_mpi_omp_shm: final/*.c compiled against the
fork/shared-memory MPI shim oracle/mpi_shm/mpi.h) on the box's host cores in several tasks x threads splits, the way the
report ran it under mpirun.  Test infrastructure: times the CPU baseline only, nothing of the product runs.
Prints the reference's own CSV line (final/SpGEMM_mpi_omp.c:336) per split and the IP/s it corresponds to."""
import importlib, os, subprocess, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench
bs = importlib.import_module("binary-spgemm_b200")
w = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
times = int(sys.argv[2]) if len(sys.argv) > 2 else 3
gname, gargs, desc = bench.WORKLOADS[w]
row, col = getattr(bs, gname)(*gargs)
n = len(row) - 1
ip = int(np.diff(row).astype(np.int64)[col].sum())
path = f"/dev/shm/ref_{w}.mtx"
t0 = time.perf_counter(); bs.write_mtx(path, row, col); print(f"# {desc}: n={n} nnz={len(col)} IP={ip}; wrote {path} in {time.perf_counter()-t0:.1f} s", flush=True)
exe = ROOT / "oracle" / "_ref" / "SpGEMM_mpi_omp_shm"
ncpu = os.cpu_count() or 1
C = 1
while C * 2 <= ncpu: C *= 2
splits = sorted({(1, C), (C, 1)} | {(p, C // p) for p in (2, 4, 8) if p < C})
if len(sys.argv) > 3:
    splits = [tuple(int(x) for x in t.split("x")) for t in sys.argv[3].split(",")]
print(f"# host cpus {ncpu}; splits (tasks x threads): {splits}")
for P, T in splits:
    block = n // (P * T)
    env = dict(os.environ, MPI_SHIM_TASKS=str(P), OMP_NUM_THREADS=str(T), OMP_STACKSIZE="1G")
    t0 = time.perf_counter()
    r = subprocess.run(["bash", "-c", f"ulimit -s unlimited 2>/dev/null; exec {exe} {path} {block} {T} {times}"], env=env, capture_output=True, text=True, timeout=1200)
    wall = time.perf_counter() - t0
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else f"(no output, exit {r.returncode}: {r.stderr[-200:]})"
    try:
        f = line.split(",")
        fastest, mean = float(f[10]), float(f[8])
        print(f"{P:3d} x {T:3d} x {block:8d}: {line}   -> {ip/mean:.3e} IP/s (mean), {ip/fastest:.3e} IP/s (fastest); wall {wall:.1f} s incl. {P}x file parse", flush=True)
    except Exception:
        print(f"{P:3d} x {T:3d}: {line}", flush=True)
os.remove(path)
