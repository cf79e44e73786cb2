"""The C++ caller of the C ABI (integration/smoke_main.cpp: no Python, no torch between the caller and libqmcb.so) on
the GPU, and a spline table restored from a QMCPACK-format dump evaluated against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from qmcpack_b200 import api as a, build
    build.build()
    a.init(0)
    return a
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_cpp_harness_runs_on_the_gpu():
    exe = os.path.join(ROOT, "integration", "_build", "qmcb_smoke")
    if not os.path.exists(exe):
        r = subprocess.run([os.path.join(ROOT, "integration", "check.sh")], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "qmcb_smoke ok" in p.stdout, (p.returncode, p.stdout, p.stderr)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_sposet_from_spline_dump_matches_oracle(api, orc, tmp_path, dt):
    from qmcpack_b200.mini_hdf5 import write_h5
    from qmcpack_b200.spline_dump import sposet_from_dump
    from qmcpack_b200.workload import random_table
    norb = 40
    coefs = random_table((7, 6, 8), norb, dt, seed=3)
    p = str(tmp_path / "einspline.tile_100010001.spin_0.tw_0.l0u40.g7x6x8.h5")
    write_h5(p, {"class_name": "SplineR2R", "sizeof": np.array(np.dtype(dt).itemsize, np.int32), "spline_0": coefs})
    lat = np.array([[5.0, 0.3, 0.0], [0.1, 4.5, 0.2], [0.0, 0.4, 5.5]])
    G = np.linalg.inv(lat)
    spo = sposet_from_dump(p, norb, G)
    r = np.random.default_rng(5).random((13, 3)) @ lat
    psi, dpsi, d2psi = spo.mw_evaluateVGL(r)
    opsi, odpsi, od2psi = orc.r2r_vgl(coefs, G, norb, r)
    tol = 1e-5 if dt == np.float32 else 1e-10
    for a, b in ((psi, opsi), (dpsi, odpsi), (d2psi, od2psi)):
        assert np.abs(a - b).max() <= tol * np.abs(b).max()
