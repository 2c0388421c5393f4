"""oracle/oracle.py — ctypes loader for the CPU checker.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  It exposes two things:

* ``Oracle``  — liboracle.so, our C restatement of the reference's algorithm (oracle/spgemm_oracle.c);
* ``Ref``     — oracle/_ref/libref_spgemm.so, the UNMODIFIED reference (final/*.c compiled in place by
                ``make -C oracle ref``); its functions are the reference's own SpGEMM_bigslice /
                SpGEMM_omp / SpGEMM_mpi / readCOO / coo2csc (final/SpGEMM_mpi_omp.c:15-225,
                final/utils.c:47-81, final/coo2csc.c:22-64).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ORACLE_LIB = HERE / "liboracle.so"
REF_DIR = HERE / "_ref"
REF_LIB = REF_DIR / "libref_spgemm.so"

_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]
_libc.malloc.restype = C.c_void_p
_libc.malloc.argtypes = [C.c_size_t]


def build(ref: bool = True):
    """Compile liboracle.so and, when /root/reference is present, oracle/_ref (outputs only)."""
    subprocess.run(["make", "-C", str(HERE), "all"], check=True, capture_output=True)
    if ref:
        subprocess.run(["make", "-C", str(HERE), "ref"], check=True, capture_output=True)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Oracle:
    def __init__(self):
        if not ORACLE_LIB.exists():
            build(ref=False)
        L = C.CDLL(str(ORACLE_LIB))
        L.oracle_intermediate_products.restype = C.c_int64
        L.oracle_intermediate_products.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
        L.oracle_spgemm.restype = C.c_int64
        L.oracle_spgemm.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                    C.POINTER(C.c_void_p), C.c_void_p, C.c_int32]
        L.oracle_spgemm_rows.restype = C.c_int64
        L.oracle_spgemm_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                         C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_coo2csc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.oracle_coo2csc.restype = None
        L.oracle_readCOO.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.oracle_max_threads.restype = C.c_int
        L.oracle_set_threads.argtypes = [C.c_int]
        self.L = L

    def max_threads(self) -> int:
        return self.L.oracle_max_threads()

    def set_threads(self, t: int):
        self.L.oracle_set_threads(t)

    def intermediate_products(self, Acol, Arow, An, Brow) -> int:
        Acol, Arow, Brow = _i32(Acol), _i32(Arow), _i32(Brow)
        return self.L.oracle_intermediate_products(Acol.ctypes.data, Arow.ctypes.data, An, Brow.ctypes.data)

    def spgemm(self, Acol, Arow, An, Bcol, Brow, Bm, nslices: int = 0):
        """-> (Ccol int32[nnz], Crow int64[An+1])"""
        Acol, Arow, Bcol, Brow = _i32(Acol), _i32(Arow), _i32(Bcol), _i32(Brow)
        Crow = np.zeros(An + 1, dtype=np.int64)
        out = C.c_void_p()
        if nslices <= 0:
            nslices = max(1, min(An, 4 * self.max_threads()))
        nnz = self.L.oracle_spgemm(Acol.ctypes.data, Arow.ctypes.data, An, Bcol.ctypes.data, Brow.ctypes.data, Bm,
                                   C.byref(out), Crow.ctypes.data, nslices)
        if nnz < 0:
            raise MemoryError("oracle_spgemm failed")
        Ccol = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_int32)), shape=(max(nnz, 1),))[:nnz].copy()
        self.L.oracle_free(out)
        return Ccol, Crow

    def spgemm_masked(self, Acol, Arow, An, Bcol, Brow, Bm, Fcol, Frow):
        """C = F .* (A·B) (SpGEMM_masked, final/SpGEMM_mpi_omp.c:232-288) -> (Ccol int32[nnz], Crow int64[An+1])"""
        Acol, Arow, Bcol, Brow, Fcol, Frow = (_i32(x) for x in (Acol, Arow, Bcol, Brow, Fcol, Frow))
        Crow = np.zeros(An + 1, dtype=np.int64)
        out = C.c_void_p()
        self.L.oracle_spgemm_masked.restype = C.c_int64
        self.L.oracle_spgemm_masked.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                                C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p]
        nnz = self.L.oracle_spgemm_masked(Acol.ctypes.data, Arow.ctypes.data, An, Bcol.ctypes.data, Brow.ctypes.data, Bm,
                                          Fcol.ctypes.data, Frow.ctypes.data, C.byref(out), Crow.ctypes.data)
        if nnz < 0:
            raise MemoryError("oracle_spgemm_masked failed")
        Ccol = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_int32)), shape=(max(nnz, 1),))[:nnz].copy()
        self.L.oracle_free(out)
        return Ccol, Crow

    def coo2csc(self, row_coo, col_coo, n, is_one_based=0):
        r = np.ascontiguousarray(row_coo, dtype=np.uint32)
        c = np.ascontiguousarray(col_coo, dtype=np.uint32)
        out_row = np.empty(max(len(r), 1), dtype=np.uint32)
        out_col = np.empty(n + 1, dtype=np.uint32)
        self.L.oracle_coo2csc(out_row.ctypes.data, out_col.ctypes.data, r.ctypes.data, c.ctypes.data, len(r), n, is_one_based)
        return out_row[: len(r)], out_col

    def readCOO(self, path):
        row, col = C.c_void_p(), C.c_void_p()
        M, N, nnz = C.c_uint32(), C.c_uint32(), C.c_uint32()
        rc = self.L.oracle_readCOO(os.fsencode(str(path)), C.byref(row), C.byref(col), C.byref(M), C.byref(N), C.byref(nnz))
        if rc:
            raise OSError(f"oracle_readCOO failed: {rc}")
        r = np.ctypeslib.as_array(C.cast(row, C.POINTER(C.c_uint32)), shape=(M.value + 1,)).copy()
        c = np.ctypeslib.as_array(C.cast(col, C.POINTER(C.c_uint32)), shape=(max(nnz.value, 1),))[: nnz.value].copy()
        self.L.oracle_free(row)
        self.L.oracle_free(col)
        return r, c, M.value, N.value, nnz.value


class Ref:
    """The reference's own compiled functions (oracle/_ref/libref_spgemm.so).  32-bit `int` everywhere."""

    def __init__(self):
        if not REF_LIB.exists():
            raise FileNotFoundError(f"{REF_LIB} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(str(REF_LIB))
        sig = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.SpGEMM_bigslice.argtypes = sig + [C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int]
        L.SpGEMM_bigslice.restype = None
        L.SpGEMM_omp.argtypes = sig + [C.POINTER(C.c_void_p), C.c_void_p, C.c_int]
        L.SpGEMM_omp.restype = None
        L.SpGEMM_mpi.argtypes = sig + [C.POINTER(C.c_void_p), C.c_void_p, C.c_int]
        L.SpGEMM_mpi.restype = None
        L.readCOO.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                              C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.readCOO.restype = None
        L.coo2csc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.coo2csc.restype = None
        self.L = L
        self._gomp = None

    def set_threads(self, t: int):
        """omp_set_num_threads, as the reference driver does (final/SpGEMM_mpi_omp.c:305)."""
        if self._gomp is None:
            self._gomp = C.CDLL("libgomp.so.1")
            self._gomp.omp_set_num_threads.argtypes = [C.c_int]
        self._gomp.omp_set_num_threads(int(t))

    def max_threads(self) -> int:
        if self._gomp is None:
            self._gomp = C.CDLL("libgomp.so.1")
            self._gomp.omp_set_num_threads.argtypes = [C.c_int]
        self._gomp.omp_get_max_threads.restype = C.c_int
        return self._gomp.omp_get_max_threads()

    def bigslice(self, Acol, Arow, An, Bcol, Brow, Bm, start_row=0, end_row=None):
        """SpGEMM_bigslice (final/SpGEMM_mpi_omp.c:15-58) -> (Ccol, slice-relative Crow)."""
        Acol, Arow, Bcol, Brow = _i32(Acol), _i32(Arow), _i32(Bcol), _i32(Brow)
        end_row = An if end_row is None else end_row
        Crow = np.zeros(end_row - start_row + 1, dtype=np.int32)
        size = C.c_int(max(Bm, 1))
        buf = C.c_void_p(_libc.malloc(max(Bm, 1) * 4))     # the reference pre-sizes Ccol to Bm ints (:85-90)
        self.L.SpGEMM_bigslice(Acol.ctypes.data, Arow.ctypes.data, An, Bcol.ctypes.data, Brow.ctypes.data, Bm,
                               C.byref(buf), Crow.ctypes.data, C.byref(size), start_row, end_row)
        nnz = int(Crow[-1])
        Ccol = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_int32)), shape=(max(nnz, 1),))[:nnz].copy()
        _libc.free(buf)
        return Ccol, Crow

    def _whole(self, fn, Acol, Arow, An, Bcol, Brow, Bm, tBlock):
        Acol, Arow, Bcol, Brow = _i32(Acol), _i32(Arow), _i32(Bcol), _i32(Brow)
        Crow = np.zeros(An + 1, dtype=np.int32)
        out = C.c_void_p()
        fn(Acol.ctypes.data, Arow.ctypes.data, An, Bcol.ctypes.data, Brow.ctypes.data, Bm, C.byref(out), Crow.ctypes.data, tBlock)
        nnz = int(Crow[An])
        Ccol = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_int32)), shape=(max(nnz, 1),))[:nnz].copy()
        _libc.free(out)
        return Ccol, Crow

    def masked(self, Acol, Arow, An, Bcol, Brow, Bm, Fcol, Frow):
        """The reference's own SpGEMM_masked (final/SpGEMM_mpi_omp.c:232-288; defined there, never called by its drivers).
        Square matrices only: its flag array has An entries (:239)."""
        Acol, Arow, Bcol, Brow, Fcol, Frow = (_i32(x) for x in (Acol, Arow, Bcol, Brow, Fcol, Frow))
        self.L.SpGEMM_masked.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(C.c_int)]
        self.L.SpGEMM_masked.restype = None
        Crow = np.zeros(An + 1, dtype=np.int32)
        size = C.c_int(max(An, 1))
        buf = C.c_void_p(_libc.malloc(max(An, 1) * 4))
        self.L.SpGEMM_masked(Acol.ctypes.data, Arow.ctypes.data, An, Bcol.ctypes.data, Brow.ctypes.data, Bm,
                             Fcol.ctypes.data, Frow.ctypes.data, C.byref(buf), Crow.ctypes.data, C.byref(size))
        nnz = int(Crow[An])
        Ccol = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_int32)), shape=(max(nnz, 1),))[:nnz].copy()
        _libc.free(buf)
        return Ccol, Crow

    def omp(self, Acol, Arow, An, Bcol, Brow, Bm, tBlock):
        """SpGEMM_omp (final/SpGEMM_mpi_omp.c:71-143); An must be divisible by tBlock (:77)."""
        return self._whole(self.L.SpGEMM_omp, Acol, Arow, An, Bcol, Brow, Bm, tBlock)

    def mpi(self, Acol, Arow, An, Bcol, Brow, Bm, tBlock):
        """SpGEMM_mpi (final/SpGEMM_mpi_omp.c:155-225) with the single-rank shim."""
        return self._whole(self.L.SpGEMM_mpi, Acol, Arow, An, Bcol, Brow, Bm, tBlock)

    def omp_timed(self, Acol, Arow, An, Bcol, Brow, Bm, tBlock):
        """Runs SpGEMM_omp on pre-converted arrays, frees the result, returns (seconds, nnz).  For the CPU baseline."""
        import time
        Crow = np.zeros(An + 1, dtype=np.int32)
        out = C.c_void_p()
        t0 = time.perf_counter()
        self.L.SpGEMM_omp(Acol.ctypes.data, Arow.ctypes.data, An, Bcol.ctypes.data, Brow.ctypes.data, Bm,
                          C.byref(out), Crow.ctypes.data, tBlock)
        dt = time.perf_counter() - t0
        _libc.free(out)
        return dt, int(Crow[An])

    def readCOO(self, path):
        row, col = C.c_void_p(), C.c_void_p()
        M, N, nnz = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self.L.readCOO(os.fsencode(str(path)), C.byref(row), C.byref(col), C.byref(M), C.byref(N), C.byref(nnz))
        r = np.ctypeslib.as_array(C.cast(row, C.POINTER(C.c_uint32)), shape=(M.value + 1,)).copy()
        c = np.ctypeslib.as_array(C.cast(col, C.POINTER(C.c_uint32)), shape=(max(nnz.value, 1),))[: nnz.value].copy()
        _libc.free(row)
        _libc.free(col)
        return r, c, M.value, N.value, nnz.value

    def coo2csc(self, row_coo, col_coo, n, is_one_based=0):
        r = np.ascontiguousarray(row_coo, dtype=np.uint32)
        c = np.ascontiguousarray(col_coo, dtype=np.uint32)
        out_row = np.empty(max(len(r), 1), dtype=np.uint32)
        out_col = np.empty(n + 1, dtype=np.uint32)
        self.L.coo2csc(out_row.ctypes.data, out_col.ctypes.data, r.ctypes.data, c.ctypes.data, len(r), n, is_one_based)
        return out_row[: len(r)], out_col


def have_ref() -> bool:
    return REF_LIB.exists()
