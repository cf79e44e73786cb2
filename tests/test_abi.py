"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/qmcb.h declares,
and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from qmcpack_b200 import build, api
    build.build()
    return api.lib()


def header_symbols(name="qmcb.h"):
    txt = open(os.path.join(ROOT, "include", name)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(qmcb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(lib):
    from qmcpack_b200 import api
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/qmcb.h but not exported by libqmcb.so"
    assert sorted(api.SYMBOLS) == syms
    dsyms = header_symbols("qmcb_driver.h")
    for s in dsyms:
        assert hasattr(lib, s), f"{s} declared in include/qmcb_driver.h but not exported"
    assert sorted(api.DRIVER_SYMBOLS) == dsyms


def test_aligned_size_matches_reference_rule(lib):
    # getAlignedSize<T>: round up to 64 bytes (Platforms/CPU/SIMD/aligned_allocator.hpp:41-47)
    assert lib.qmcb_aligned_size(1, 192) == 192
    assert lib.qmcb_aligned_size(1, 193) == 208
    assert lib.qmcb_aligned_size(0, 4) == 8
    assert lib.qmcb_aligned_size(0, 385) == 392


def test_no_cpu_fallback(lib):
    """Without a CUDA device every compute entry point must fail loudly instead of falling back."""
    from qmcpack_b200 import api
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    coefs = np.zeros((4, 4, 4, 16), np.float32)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        api.SplineSPOSet(coefs, 4, np.eye(3))
    with pytest.raises(RuntimeError):
        api.init(0)


def test_product_does_not_import_oracle():
    """The product package must never route through the oracle (tests/, smoke() and bench baselines only)."""
    pkg = os.path.join(ROOT, "qmcpack_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in txt and "liboracle" not in txt and "libqmcref" not in txt, f
