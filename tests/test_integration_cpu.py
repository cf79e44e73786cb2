"""Reference-side binding (integration/): the adapters compile against the reference's own headers, the C++ caller of the
C ABI builds, links and fails loudly without a GPU, and the spline-dump reader restores a table bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
HAVE_REF = os.path.isdir("/root/reference/src")


def _run_check():
    return subprocess.run([os.path.join(ROOT, "integration", "check.sh")], capture_output=True, text=True, timeout=600)


@pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")
def test_adapters_compile_against_reference_headers():
    r = _run_check()
    assert r.returncode == 0, r.stdout + r.stderr
    assert "engine concept check: ok" in r.stdout and "spline adapter check: ok" in r.stdout


def test_cpp_harness_links_and_has_no_cpu_fallback():
    import torch
    if not os.path.exists(os.path.join(ROOT, "qmcpack_b200", "libqmcb.so")):
        pytest.skip("libqmcb.so not built")
    r = _run_check()
    assert r.returncode == 0, r.stdout + r.stderr
    exe = os.path.join(ROOT, "integration", "_build", "qmcb_smoke")
    assert os.path.exists(exe)
    if torch.cuda.is_available():
        pytest.skip("GPU present: the harness is run by tests/test_integration_gpu.py")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 2 and "no CPU fallback" in p.stderr, (p.returncode, p.stderr)


@pytest.mark.parametrize("dt,cls", [(np.float32, "SplineR2R"), (np.float64, "SplineC2COMPTarget")])
def test_spline_dump_round_trip(tmp_path, dt, cls):
    """a dump with the datasets SplineSetReader writes (class_name, sizeof, spline_0[, spline_0spline_1]) comes back bit
    for bit, blocks in order; malformed files are refused"""
    from qmcpack_b200.mini_hdf5 import write_h5
    from qmcpack_b200.spline_dump import read_spline_dump
    rng = np.random.default_rng(1)
    a = rng.normal(size=(9, 8, 10, 16)).astype(dt)
    b = rng.normal(size=(9, 8, 10, 16)).astype(dt)
    p = str(tmp_path / "dump.h5")
    write_h5(p, {"class_name": cls, "sizeof": np.array(np.dtype(dt).itemsize, np.int32), "spline_0": a,
                 "spline_0spline_1": b})
    d = read_spline_dump(p)
    assert d["class_name"] == cls and d["sizeof"] == np.dtype(dt).itemsize and d["grid"] == (6, 5, 7)
    assert d["is_complex"] == cls.startswith("SplineC2")
    assert len(d["blocks"]) == 2 and np.array_equal(d["blocks"][0], a) and np.array_equal(d["blocks"][1], b)
    write_h5(p, {"sizeof": np.array(8, np.int32), "spline_0": a.astype(np.float32)})
    with pytest.raises(ValueError):
        read_spline_dump(p)
    write_h5(p, {"something": a})
    with pytest.raises(ValueError):
        read_spline_dump(p)
