"""The C-ABI libraries load on a CPU-only host and export every symbol include/*.h declares; compute entry
points refuse to run without a GPU instead of falling back.  CPU only (no compute calls)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared(header):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(bspgemm_\w+|bs_\w+|readCOO\w*|coo2csc|tictoc)\s*\(", text))
    return names


def test_bspgemm_h_symbols_exported(bs):
    L = C.CDLL(str(bs.LIB_PATH))
    declared = _declared("bspgemm.h")
    assert declared == set(bs.ABI_SYMBOLS), declared ^ set(bs.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name


def test_bspgemm_host_h_symbols_exported(bs):
    H = C.CDLL(str(bs.HOST_LIB_PATH))
    for name in _declared("bspgemm_host.h") | set(bs.HOST_SYMBOLS):
        assert hasattr(H, name), name


def test_version_and_strerror(bs):
    L = bs.lib()
    assert b"sm_100a" in L.bspgemm_version()
    assert L.bspgemm_strerror(bs.ERR_NOGPU) and b"fallback" in L.bspgemm_strerror(bs.ERR_NOGPU)
    assert L.bspgemm_num_gpus() >= 0          # 0 before bspgemm_init (another test of this session may have initialised a context)


def test_no_cpu_fallback_without_gpu(bs):
    """On a host without CUDA devices the operators fail loudly with BSPGEMM_ERR_NOGPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    with pytest.raises(bs.BSpGEMMError) as e:
        bs.spgemm_csr([0], [0, 1], 1, [0], [0, 1], 1, 1)
    assert e.value.status == bs.ERR_NOGPU
    with pytest.raises(bs.BSpGEMMError):
        bs.DeviceSpGEMM(0)
    with pytest.raises(bs.BSpGEMMError) as e:
        bs.coo2csc_gpu([0, 1], [1, 0], 2, 0)          # the GPU converter does not fall back to the host coo2csc either
    assert e.value.status == bs.ERR_NOGPU


def test_product_does_not_link_oracle(bs):
    """The product libraries must not reference the oracle (the oracle is test infrastructure)."""
    import subprocess
    for p in (bs.LIB_PATH, bs.HOST_LIB_PATH):
        out = subprocess.run(["nm", "-D", str(p)], capture_output=True, text=True).stdout
        assert "oracle_" not in out
        ldd = subprocess.run(["ldd", str(p)], capture_output=True, text=True).stdout
        assert "liboracle" not in ldd and "libref_spgemm" not in ldd
