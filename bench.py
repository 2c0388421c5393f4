#!/usr/bin/env python
"""bench.py — boolean SpGEMM throughput on B200 (BASELINE.json metric: intermediate products / second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, NCCL)

A "step" is one pass of the hot path over the workload: C = A·A for the synthetic matrix (the reference's
drivers multiply A by itself, final/SpGEMM_mpi_omp.c:322).  Default workload = BASELINE config 3, the one
the north-star target is quoted on: uniform random boolean n=2^22, d=16, seed 1.

  value      whole-job IP/s, A and B already resident in HBM (device-resident C-ABI operator
             bspgemm_dev_multiply, B made resident once with bspgemm_dev_prepare_b — the reference replicates B
             once before its timed loop, :309 vs :318-328), CUDA events on the launching stream, max over ranks.
             `unprepared` repeats the measurement without prepare_b (B's re-layout and the probes inside every
             step).  C's device arena is grown by the first product and reused: `cold_first_call_ms` is that
             first product including its cudaMalloc.
  e2e        same metric through the host-pointer C-ABI operator bspgemm_csr_into with pinned host
             buffers: H2D of A + B and D2H of Crow + Ccol inside the timed region.  N > 1: ONE process (rank 0)
             drives all N GPUs through bspgemm_init(N) — the library's drop-in for SpGEMM_mpi: B is uploaded
             once and replicated by ncclBroadcast — while the other ranks idle; `e2e_per_rank` is the old
             figure (every rank its own single-GPU context and its own upload of B).
  validated  after the timed region every rank compares three contiguous row blocks (start / middle / end of
             its shard) of the device result with the oracle, all-reduced; a mismatch exits non-zero.
  roofline   dominant kernel (the fused symbolic+scan+fill kernel): algorithmic bytes /
             its CUDA-event duration, against the measured copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the reference's own SpGEMM_omp (oracle/_ref, compiled unmodified) on the host cores, on a
             bounded row-block sample of the same workload (rank 0, N=1 only).
Multi-GPU: A is sharded into contiguous row blocks (the reference's rank split, :165-171); B is replicated
by an NCCL broadcast (untimed, like the reference's per-rank file read :309); no data-path collective.
Total work is fixed as N grows, so scaling = "strong".
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (generator, args, description)
    "cfg2": ("gen_uniform", ((1 << 20), 8, 1), "uniform random boolean n=2^20 d=8 seed=1, C=A*A"),
    "cfg3": ("gen_uniform", ((1 << 22), 16, 1), "uniform random boolean n=2^22 d=16 seed=1, C=A*A"),
    "rmat20": ("gen_rmat", (20, 16, 0.45, 0.22, 0.22, 1), "R-MAT (.45,.22,.22,.11) scale 20 edge factor 16, C=A*A"),
    "banded22": ("gen_banded", ((1 << 22), 32), "banded n=2^22 d=32, C=A*A"),
    "cfg4": ("gen_rmat", (22, 16, 0.45, 0.22, 0.22, 1), "BASELINE config 4: R-MAT (.45,.22,.22,.11) scale 22 edge factor 16, C=A*A (int64 row pointers)"),
    "cfg5": ("gen_banded", ((1 << 24), 32), "BASELINE config 5: banded n=2^24 d=32, C=A*A"),
    # the one configuration with a PUBLISHED number: the report's n5e6_d5 (sprand, n = 5e6, nnz ~ 2.5e7), A*A in 0.62 s with
    # 20 MPI tasks on one cluster node (PDF p.3 Fig 8, read off the chart)
    "pub_n5e6_d5": ("gen_sprand", (5000000, 5.0, 1), "the report's n5e6_d5: sprand-like (Poisson(5) row lengths) n=5e6 nnz~2.5e7, C=A*A"),
    "small": ("gen_uniform", ((1 << 16), 8, 1), "uniform random boolean n=2^16 d=8 seed=1, C=A*A"),
}


# ------------------------------------------------------------------------------------------------ sharding helpers (also used by tests/test_multirank.py)
def shard_bounds(n: int, rank: int, world: int):
    """Contiguous row block of `rank`: [n*rank/world, n*(rank+1)/world) — the reference's tasksize split
    (final/SpGEMM_mpi_omp.c:165,171) generalised to any n."""
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_bounds_ip(row_t, col_t, n: int, world: int):
    """Contiguous row blocks of equal WORK: split points on the prefix sum of the rows' intermediate products
    (+1 per row).  row_t / col_t: the CSR of A = B as tensors (any device).  Returns the world+1 boundaries; the product is the
    same whatever the split."""
    import torch
    blen = (row_t[1:] - row_t[:-1]).to(torch.int64)
    cs = torch.zeros(col_t.numel() + 1, dtype=torch.int64, device=row_t.device)
    cs[1:] = torch.cumsum(blen[col_t.long()], 0)                  # prefix of "length of the selected B row" over A's nonzeros
    r = row_t.long()
    row_ip = cs[r[1:]] - cs[r[:-1]]                               # intermediate products of every row
    pre = torch.zeros(n + 1, dtype=torch.int64, device=row_t.device)
    pre[1:] = torch.cumsum(row_ip + 1, 0)                         # (+1: an empty row still costs a row)
    targets = torch.tensor([int(pre[-1]) * q // world for q in range(1, world)], dtype=torch.int64, device=row_t.device)
    cuts = torch.searchsorted(pre, targets).clamp(max=n).tolist()
    b = [0] + cuts + [n]
    for q in range(1, len(b)):
        b[q] = max(b[q], b[q - 1])
    return b


def broadcast_csr(row, col, n: int, src: int, device):
    """Replicate a CSR matrix from `src` to every rank (torch.distributed broadcast: NCCL over NVLink on GPUs,
    gloo on CPU).  Stands in for every MPI rank reading the whole file (final/SpGEMM_mpi_omp.c:309)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    meta = torch.zeros(1, dtype=torch.int64, device=device)
    if rank == src:
        meta[0] = len(col)
    dist.broadcast(meta, src)
    nnz = int(meta.item())
    if rank == src:
        row_t = torch.from_numpy(np.ascontiguousarray(row, dtype=np.int32)).to(device)
        col_t = torch.from_numpy(np.ascontiguousarray(col, dtype=np.int32)).to(device)
    else:
        row_t = torch.empty(n + 1, dtype=torch.int32, device=device)
        col_t = torch.empty(nnz, dtype=torch.int32, device=device)
    dist.broadcast(row_t, src)
    dist.broadcast(col_t, src)
    return row_t, col_t


def gather_shards(Ccol_t, Crow_t, n: int, rank: int, world: int, device):
    """Validation-only gather of the per-rank CSR slices to rank 0 (replaces MPI_Gather/Gatherv + the root
    fix-up loop, final/SpGEMM_mpi_omp.c:178-223).  Crow_t is slice-relative.  Returns (Ccol, Crow) on rank 0."""
    import torch
    import torch.distributed as dist
    cnt = torch.tensor([int(Crow_t[-1])], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(counts, cnt)
    counts = [int(c.item()) for c in counts]
    disp = np.concatenate([[0], np.cumsum(counts)])
    if rank != 0:
        dist.send(Crow_t.to(torch.int64).contiguous(), 0)
        if counts[rank]:
            dist.send(Ccol_t.contiguous(), 0)
        return None
    Crow = torch.zeros(n + 1, dtype=torch.int64, device=device)
    Ccol = torch.empty(int(disp[-1]), dtype=torch.int32, device=device)
    for r in range(world):
        r0, r1 = shard_bounds(n, r, world)
        if r == 0:
            cr, cc = Crow_t.to(torch.int64), Ccol_t
        else:
            cr = torch.empty(r1 - r0 + 1, dtype=torch.int64, device=device)
            dist.recv(cr, r)
            cc = torch.empty(counts[r], dtype=torch.int32, device=device)
            if counts[r]:
                dist.recv(cc, r)
        Crow[r0 + 1:r1 + 1] = cr[1:] + int(disp[r])
        Ccol[int(disp[r]):int(disp[r + 1])] = cc
    return Ccol, Crow


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples nvidia-smi during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.lines, self.proc = gpu_index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference_run(row, col, n, target_seconds: float, steps: int = 1, warmup: int = 0):
    """Times the reference's own SpGEMM_omp (final/SpGEMM_mpi_omp.c:71-143, compiled unmodified into
    oracle/_ref) — or the oracle port when that build is absent — on a bounded contiguous row block of A
    against the full B, with every host thread it can use.  Returns a dict for the JSON line."""
    from oracle.oracle import Oracle, Ref, have_ref     # the one place bench.py may execute oracle/
    ncpu = os.cpu_count() or 1
    T = 1
    while T * 2 <= ncpu:
        T *= 2
    kind = "reference" if have_ref() else "port"
    eng = Ref() if kind == "reference" else Oracle()
    eng.set_threads(T)
    row32, col32 = np.ascontiguousarray(row, np.int32), np.ascontiguousarray(col, np.int32)
    blen = np.diff(row32).astype(np.int64)

    def run(rows):
        rows = max(T, (rows // T) * T)                       # the reference needs rows % tBlock == 0 (:77)
        ip = int(blen[col32[row32[0]:row32[rows]]].sum())
        if kind == "reference":
            dt, nnz = eng.omp_timed(col32, row32, rows, col32, row32, n, rows // T)
        else:
            t0 = time.perf_counter()
            cc, cr = eng.spgemm(col32, row32, rows, col32, row32, n, nslices=T)
            dt, nnz = time.perf_counter() - t0, int(cr[-1])
        return rows, ip, dt, nnz

    # calibrate on a small block, then size the sample for ~target_seconds per step
    rows, ip, dt, _ = run(min(n, max(T, 1 << 14)))
    rate = ip / max(dt, 1e-6)
    want_ip = rate * target_seconds
    avg_ip_per_row = max(1.0, ip / rows)
    rows = int(min(n, max(T, want_ip / avg_ip_per_row)))
    times, ips, nnzs = [], [], []
    for s in range(warmup + steps):
        r, ip, dt, nnz = run(rows)
        if s >= warmup:
            times.append(dt); ips.append(ip); nnzs.append(nnz)
    mean_t = float(np.mean(times))
    return {
        "value": float(np.mean(ips) / mean_t), "unit": "IP/s", "cores": T, "kind": kind,
        "sample": f"rows [0,{r}) of A ({r}/{n} rows, {ips[-1]} IP, {nnzs[-1]} output nnz) x full B; "
                  f"tasks x threads x block = 1 x {T} x {r // T}; {mean_t:.3f} s/step",
        "ms_per_step": mean_t * 1e3, "out_nnz_per_s": float(np.mean(nnzs) / mean_t), "host_cpus": ncpu,
    }


# ------------------------------------------------------------------------------------------------ main
_RESULT_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version line on the GPU
    boxes): keep a private copy of the real stdout for the result line and point fd 1 at stderr for everything else."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "fused", "twophase"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true", help="development runs: skip the host-pointer end-to-end measurement (e2e: null)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="target CPU work per reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--validate", action="store_true", help="gather shards and compare with the oracle (small workloads)")
    ap.add_argument("--validate-rows", type=int, default=21000, help="rows per rank compared with the oracle after the timed region (0 = off)")
    ap.add_argument("--no-prepare", action="store_true", help="headline value without bspgemm_dev_prepare_b")
    ap.add_argument("--split", default="rows", choices=["rows", "ip"], help="row blocks of equal rows (the reference's split) or of equal intermediate products")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    gname, gargs, gdesc = WORKLOADS[args.workload]
    n = (1 << gargs[0]) if gname == "gen_rmat" else gargs[0]
    config = {"workload": args.workload, "description": gdesc, "n": n, "product": "C=A*A (boolean CSR)",
              "sharding": f"{world} contiguous row block(s) of A, B replicated", "l2_policy": "inputs_larger_than_l2"}

    bs = importlib.import_module("binary-spgemm_b200")

    # ---------------- reference arm: the reference's CPU implementation on the host cores
    if args.impl == "reference":
        if rank != 0:
            return 0
        row, col = getattr(bs, gname)(*gargs)
        res = cpu_reference_run(row, col, n, args.cpu_seconds, steps=max(1, args.steps), warmup=min(args.warmup, 1))
        line = {"impl": "reference", "metric": "intermediate_products_per_sec", "value": res["value"], "unit": "IP/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": "IP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "out_nnz_per_s": res["out_nnz_per_s"], "gpu_launches": 0}
        _emit(line)
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- workload: generated on rank 0's host, replicated over NCCL (untimed)
    t0 = time.time()
    if rank == 0:
        row, col = getattr(bs, gname)(*gargs)
    else:
        row = col = None
    if world > 1:
        d_row, d_col = broadcast_csr(row, col, n, 0, dev)
    else:
        d_row, d_col = torch.from_numpy(row).to(dev), torch.from_numpy(col).to(dev)
    nnzA = int(d_col.numel())
    if args.split == "ip" and world > 1:
        bounds = shard_bounds_ip(d_row, d_col, n, world)
        r0, r1 = bounds[rank], bounds[rank + 1]
        config["sharding"] = f"{world} contiguous row blocks of A with equal intermediate products (bounds {bounds}), B replicated"
        os.environ["BSPGEMM_SPLIT"] = "ip"              # the single-process N-GPU operator (e2e) splits the same way
    else:
        r0, r1 = shard_bounds(n, rank, world)
    rows = r1 - r0
    shard_nnz = int(d_row[r1].item()) - int(d_row[r0].item())
    gen_s = time.time() - t0

    mode = {"auto": bs.MODE_AUTO, "fused": bs.MODE_FUSED, "twophase": bs.MODE_TWOPHASE}[args.mode]
    h = bs.DeviceSpGEMM(local_rank, mode)
    i64 = args.workload.startswith("rmat") or args.workload == "cfg4"          # nnz(C) of the R-MAT workloads exceeds 2^31: 64-bit row pointers
    d_crow = torch.zeros(rows + 1, dtype=torch.int64 if i64 else torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    a_row_ptr = d_row.data_ptr() + 4 * r0             # shifted Arow, absolute offsets (final/SpGEMM_mpi_omp.c:171)

    # the product call with its arguments marshalled once (the C call is what a C host would issue in its timing loop)
    step = h.bound_multiply(d_col, a_row_ptr, rows, shard_nnz, d_col, d_row, n, n, nnzA, d_crow, stream=stream.cuda_stream, crow_is_i64=i64)

    def timed(K):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        stats = []
        barrier()
        w0 = time.perf_counter()
        for k in range(K):
            ev[k][0].record(stream)
            out = step()
            ev[k][1].record(stream)
            stats.append(h.stats())
        barrier()
        wall = time.perf_counter() - w0
        return [a.elapsed_time(b) for a, b in ev], stats, wall, out

    # cold first call: the handle's arenas do not exist yet (device allocation of C inside, like the reference's mallocs :88-92,115)
    barrier()
    w0 = time.perf_counter()
    ptr, nnz = step()
    torch.cuda.synchronize()
    cold_ms = (time.perf_counter() - w0) * 1e3
    # ---------------- without prepare_b: B's re-layout and the probes inside every step
    for _ in range(max(3, args.warmup)):
        ptr, nnz = step()
    un_ms, un_stats, _, _ = timed(min(args.steps, 5))
    un_t = torch.tensor([sum(un_ms) / len(un_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(un_t, op=dist.ReduceOp.MAX)
    # ---------------- B resident once (untimed, like the reference's per-rank file read), then the timed region:
    #                  K steps, CUDA events on the launching stream
    if not args.no_prepare:
        h.prepare_b(d_col, d_row, n, n, nnzA, stream=stream.cuda_stream)
    for _ in range(max(3, args.warmup)):
        ptr, nnz = step()
    with ClockSampler(local_rank) as clk:
        dev_ms, stats, wall, (ptr, nnz) = timed(args.steps)
    clocks = clk.summary()
    t_local = sum(dev_ms) / 1e3
    tt = torch.tensor([t_local, wall], dtype=torch.float64, device=dev)
    agg = torch.tensor([float(stats[-1]["ip"]), float(nnz), float(stats[-1]["algorithmic_bytes"]), float(sum(s["launches"] for s in stats))],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    t_max, wall_max = float(tt[0]), float(tt[1])
    ip_total, nnz_total, alg_bytes_total, launches_total = (float(x) for x in agg)
    ms_per_step = t_max / args.steps * 1e3

    # ---------------- roofline of the dominant kernel (this rank's launch; rank 0 reports)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    main_ms = float(np.mean([s["ms_main"] for s in stats]))
    est_ms = float(np.mean([s["ms_estimate"] for s in stats]))
    alg_bytes_launch = float(stats[-1]["algorithmic_bytes"])
    achieved = alg_bytes_launch / (main_ms * 1e-3) / 1e9 if main_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.loads((ROOT / "profiles" / "traffic.json").read_text()).get(f"{args.workload}@{world}")   # ncu dram bytes of THIS workload at THIS N, else null
    except Exception:
        pass
    last = stats[-1]
    if last["mode"] != bs.MODE_FUSED:
        kernel_name = "scan + k_rows_warp<G,MODE_FILL>"
    elif last.get("variant", 0) == 2:
        kernel_name = "k_fused_sort%s<W=%d> (%d rows per tile)" % ("_async" if last.get("kernel_flags", 0) & 1 else "", last["group"], last["rows_per_tile"])
    elif last.get("variant", 0) == 3:
        kernel_name = "k_band (%d rows per tile, 128-bit register bitmap per row)" % last["rows_per_tile"]
    elif last.get("variant", 0) == 1:
        kernel_name = "k_fused_ell<W=%d,R=%d>" % (last["group"], last["rows_per_tile"])
    elif last.get("kernel_flags", 0) & 4:
        kernel_name = "k_rows_tiny + k_rows_warp<G=%d> count, k_scan, fill (small rows in two unordered passes)" % last["group"]
    else:
        kernel_name = "k_fused<G=%d>" % last["group"]
    roof_ms = main_ms
    if main_ms < 0.5 * ms_per_step:
        # skewed matrices: the step is several kernels of comparable size (big-row kernels, copy, small rows) and the fused / fill
        # kernel is a small part of it — the roofline is then taken over the whole step, which is what the algorithmic bytes describe
        sym_ms, num_ms = float(np.mean([s["ms_symbolic"] for s in stats])), float(np.mean([s["ms_numeric"] for s in stats]))
        kernel_name = "whole step: big-row kernels k_rows_bm + k_rows_sort<8,256> %.1f ms, %s %.1f ms, k_copy_rows %.1f ms" % (sym_ms, kernel_name, main_ms, num_ms)
        roof_ms = ms_per_step
        achieved = alg_bytes_launch / (roof_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                "traffic": traffic, "kernel": kernel_name,
                "kernel_ms": roof_ms, "algorithmic_bytes_per_launch": alg_bytes_launch, "peak_source": peak_src,
                "frac_of_8TBs": achieved / 8000.0, "step_frac": (alg_bytes_launch / (ms_per_step * 1e-3) / 1e9) / peak_gbs}
    # where neighbouring rows re-gather the same B rows (banded) the real traffic is far below the
    # algorithmic figure; the compulsory one (A, B and C once) is reported beside it
    try:
        comp = 4.0 * (rows + 1) + 4.0 * float(shard_nnz) + 4.0 * (n + 1) + 4.0 * float(len(col)) + 4.0 * float(stats[-1]["nnz"]) + (8.0 if i64 else 4.0) * (rows + 1)
        roofline["compulsory_bytes_per_launch"] = comp
        roofline["frac_compulsory"] = (comp / (roof_ms * 1e-3) / 1e9) / peak_gbs if roof_ms > 0 else 0.0
    except Exception:
        pass

    # ---------------- every rank checks row blocks of its device result against the oracle (checker only, outside the timed region)
    validated = None
    if args.validate_rows > 0:
        from oracle.oracle import Oracle
        orc = Oracle()
        if rank == 0 and row is not None:
            row_h, col_h = np.ascontiguousarray(row, np.int32), np.ascontiguousarray(col, np.int32)
        else:
            row_h, col_h = d_row.cpu().numpy(), d_col.cpu().numpy()
        per = max(1, min(rows, args.validate_rows) // 3)
        starts = sorted({0, max(0, rows // 2 - per // 2), max(0, rows - per)})
        crow_dev = d_crow.to(torch.int64)
        Cview = bs.device_view(ptr, nnz, local_rank)
        ok, checked, blocks = True, 0, []
        for b0 in starts:
            b1 = min(rows, b0 + per)
            if b1 <= b0:
                continue
            wc, wr = orc.spgemm(col_h, row_h[r0 + b0:], b1 - b0, col_h, row_h, n)
            gr = crow_dev[b0:b1 + 1].cpu().numpy()
            good = bool((gr - gr[0] == wr).all()) and int(gr[-1] - gr[0]) == len(wc)
            if good:
                good = bool((Cview[int(gr[0]):int(gr[-1])].cpu().numpy() == wc).all())
            ok = ok and good
            checked += b1 - b0
            blocks.append([r0 + b0, r0 + b1])
        vt = torch.tensor([1.0 if ok else 0.0, float(checked)], dtype=torch.float64, device=dev)
        vmin = vt.clone()
        if world > 1:
            dist.all_reduce(vmin, op=dist.ReduceOp.MIN)
            dist.all_reduce(vt, op=dist.ReduceOp.SUM)
        validated = {"ok": bool(vmin[0] > 0.5), "rows": int(vt[1]), "ranks": world, "rank0_blocks": blocks,
                     "against": "oracle/spgemm_oracle.c (64-bit row pointers) on the same rows, bit-exact Crow and Ccol"}
        del Cview, crow_dev
        if not validated["ok"]:
            print(f"# rank {rank}: device result differs from the oracle (ok={ok})", file=sys.stderr)

    # ---------------- optional validation against the oracle (small workloads only)
    if args.validate:
        Ccol_t = bs.device_view(ptr, nnz, local_rank).clone()
        full = gather_shards(Ccol_t, d_crow, n, rank, world, dev) if world > 1 else (Ccol_t, d_crow.to(torch.int64))
        if rank == 0:
            from oracle.oracle import Oracle
            wc, wr = Oracle().spgemm(col, row, n, col, row, n)
            ok = bool((full[1].cpu().numpy() == wr).all() and (full[0].cpu().numpy() == wc).all())
            print(f"# validate vs oracle: {'ok' if ok else 'MISMATCH'}", file=sys.stderr)
            if not ok:
                return 3
    h.close()
    del h

    # ---------------- end to end: host CSR in -> host CSR out through the host-pointer C-ABI operator
    e2e = None
    e2e_per_rank = None
    if not args.no_e2e:
        if world > 1:
            rc = torch.empty(n + 1, dtype=torch.int32).pin_memory(); rc.copy_(d_row)
            cc = torch.empty(nnzA, dtype=torch.int32).pin_memory(); cc.copy_(d_col)
            row_h, col_h = rc.numpy(), cc.numpy()
        else:
            rc = torch.from_numpy(row).pin_memory(); cc = torch.from_numpy(col).pin_memory()
            row_h, col_h = rc.numpy(), cc.numpy()
        out_cap = 16 if i64 else int(nnz) + 16
        out_pin = torch.empty(out_cap, dtype=torch.int32).pin_memory()
        out_h = out_pin.numpy()
        torch.cuda.synchronize()
        bs.init(1, devices=[local_rank])
        e2e_steps = max(1, min(args.e2e_steps, args.steps))
        a_row_h = row_h[r0:r1 + 1]
        if i64:      # nnz(C) >= 2^31: the 64-bit row-pointer operator (callee-allocated output)
            def e2e_call():
                cc_, cr_ = bs.spgemm_csr(col_h, a_row_h, rows, col_h, row_h, n, n, i64=True)
                return len(cc_), cr_
        else:
            def e2e_call():
                return bs.spgemm_csr_into(col_h, a_row_h, rows, col_h, row_h, n, n, out_h)
        e2e_nnz, _ = e2e_call()                                                                # warm-up (allocations)
        barrier()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_nnz, crow_h = e2e_call()
        barrier()
        e2e_wall = time.perf_counter() - w0
        bs.finalize()
        h2d = 4 * (rows + 1) + 4 * shard_nnz + 4 * (n + 1) + 4 * nnzA
        d2h = 4 * rows + 4 * int(e2e_nnz)
        e2 = torch.tensor([e2e_wall, 0.0], dtype=torch.float64, device=dev)
        io = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2, op=dist.ReduceOp.MAX)
            dist.all_reduce(io, op=dist.ReduceOp.SUM)
        e2e_s = float(e2[0]) / e2e_steps
        e2e = {"value": ip_total / e2e_s, "unit": "IP/s", "h2d_bytes_per_step": int(io[0]), "d2h_bytes_per_step": int(io[1]),
               "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "api": "bspgemm_csr_i64 (pinned host CSR in, malloc'ed host CSR out)" if i64 else "bspgemm_csr_into (pinned host CSR in, pinned host CSR out)",
               "timer": "host CLOCK_MONOTONIC around the synchronous call, max over ranks"}
        e2e["bound"] = "PCIe: %.1f GB/s of host<->device copies per GPU inside the call (device work is %.1f %% of it)" % (
            (h2d + d2h) / e2e_s / 1e9, 100.0 * ms_per_step * 1e-3 / e2e_s)

        # ---------------- N > 1: the library's real N-GPU drop-in, ONE process driving all GPUs (bspgemm_init(N) -> row-block shards,
        #                  B uploaded once + ncclBroadcast, gather at displacements); the other ranks free their memory and idle on the
        #                  store (a host-side wait: an NCCL barrier would park a spinning kernel on the GPUs rank 0 is about to use)
        e2e_per_rank = None
        if world > 1:
            e2e_per_rank = e2e
            e2e_per_rank["api"] += " — every rank its own single-GPU context and its own upload of B"
            del d_row, d_col, d_crow
            torch.cuda.empty_cache()
            store = dist.distributed_c10d._get_default_store()
            barrier()
            if rank == 0:
                try:
                    out_all = torch.empty(16 if i64 else int(nnz_total) + 16, dtype=torch.int32).pin_memory()
                    out_all_h = out_all.numpy()
                    bs.init(world)
                    if i64:
                        def sp_call():
                            cc_, cr_ = bs.spgemm_csr(col_h, row_h, n, col_h, row_h, n, n, i64=True)
                            return len(cc_), cr_
                    else:
                        def sp_call():
                            return bs.spgemm_csr_into(col_h, row_h, n, col_h, row_h, n, n, out_all_h)
                    sp_nnz, _ = sp_call()                                                   # warm-up (allocations, NCCL channels)
                    w0 = time.perf_counter()
                    for _ in range(e2e_steps):
                        sp_nnz, _ = sp_call()
                    sp_s = (time.perf_counter() - w0) / e2e_steps
                    bs.finalize()
                    assert int(sp_nnz) == int(nnz_total), (sp_nnz, nnz_total)
                    e2e = {"value": ip_total / sp_s, "unit": "IP/s",
                           "h2d_bytes_per_step": int(2 * (4 * (n + 1) + 4 * nnzA)), "d2h_bytes_per_step": int(4 * n + 4 * int(sp_nnz)),
                           "ms_per_step": sp_s * 1e3, "steps": e2e_steps,
                           "api": f"bspgemm_init({world}) + " + ("bspgemm_csr_i64" if i64 else "bspgemm_csr_into") + " in ONE process: A sharded into row blocks, "
                                  "B uploaded once and replicated by ncclBroadcast over NVLink, C gathered to pinned host memory at the displacements",
                           "timer": "host CLOCK_MONOTONIC around the synchronous call (rank 0; the other ranks idle)"}
                finally:
                    store.set("bspgemm_e2e_done", "1")
            else:
                store.wait(["bspgemm_e2e_done"])

    # ---------------- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = cpu_reference_run(row, col, n, args.cpu_seconds, steps=1, warmup=0)
        cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {
            "metric": "intermediate_products_per_sec", "value": ip_total / (t_max / args.steps), "unit": "IP/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_per_rank": e2e_per_rank,
            "validated": validated, "b_prepared": not args.no_prepare,
            "unprepared": {"ms_per_step": float(un_t[0]), "value": ip_total / (float(un_t[0]) * 1e-3),
                           "kernel_ms": float(np.mean([s_["ms_main"] for s_ in un_stats])),
                           "relayout_ms": float(np.mean([s_["ms_symbolic"] for s_ in un_stats])),
                           "launches_per_step": un_stats[-1]["launches"],
                           "what": "same steps without bspgemm_dev_prepare_b: B's re-layout, the probes and their host round trip inside every step"},
            "arena": {"policy": "C's device arena (and the workspace) is grown by the first product and reused by the following ones: "
                                "device allocation is outside the timed steps", "cold_first_call_ms": cold_ms},
            "gpu_launches": int(launches_total), "clocks": clocks,
            "out_nnz_per_s": nnz_total / (t_max / args.steps), "ip": int(ip_total), "nnz_c": int(nnz_total), "nnz_a": nnzA,
            "pipeline": {"mode": "fused" if stats[-1]["mode"] == bs.MODE_FUSED else "twophase", "variant": {0: "csr", 1: "ell-hash", 2: "ell-sort", 3: "band"}.get(stats[-1].get("variant", 0)),
                         "rows_per_tile": stats[-1].get("rows_per_tile", 0), "cap_s": stats[-1]["cap_s"],
                         "group": stats[-1]["group"], "rows_s": stats[-1]["rows_s"], "rows_m": stats[-1]["rows_m"], "rows_l": stats[-1]["rows_l"],
                         "ms_estimate": est_ms, "ms_main": main_ms,
                         "ms_symbolic": float(np.mean([s["ms_symbolic"] for s in stats])),
                         "ms_numeric": float(np.mean([s["ms_numeric"] for s in stats])),
                         "launches_per_step": stats[-1]["launches"], "plan_cached": bool(stats[-1].get("plan_cached", 0))},
            "wall_ms_per_step": wall_max / args.steps * 1e3, "setup_s": gen_s,
        }
        _emit(line)
    if validated is not None and not validated["ok"]:
        if world > 1:
            dist.destroy_process_group()
        return 4
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
