"""GPU parity: multi-walker tricubic B-spline SPO evaluation through the C ABI vs the CPU oracle.

Tolerances (BASELINE.json north_star): orbital values/gradients/laplacians within 1e-5 relative in mixed precision
(float tables) and 1e-10 in full precision; "relative" is taken against the largest magnitude of the compared field
because single orbitals pass through zero.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / scale


TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-10}

LATTICES = {
    "cubic": np.eye(3) * 5.0,
    "general": np.array([[4.0, 0.3, 0.1], [0.5, 5.0, -0.2], [0.2, -0.4, 6.0]]),
}


def positions(lat, nw, seed=0):
    """inside the cell, outside it (both signs), exactly on grid planes and on the cell faces"""
    rng = np.random.default_rng(seed)
    frac = rng.random((nw, 3)) * 3.0 - 1.0
    frac[0] = [0.0, 0.0, 0.0]
    frac[1] = [1.0, 1.0, 1.0]
    frac[2] = [0.5, 0.25, 0.125]
    frac[3] = [-1e-9, 0.3, 0.7]
    return frac @ lat


@pytest.fixture(scope="module")
def api():
    from qmcpack_b200 import api as a, build
    build.build()
    a.init(0)
    return a


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("latname", ["cubic", "general"])
@pytest.mark.parametrize("grid,norb", [((7, 5, 6), 19), ((6, 6, 6), 200), ((8, 8, 8), 384), ((5, 6, 7), 193)])
def test_r2r_vgl_matches_oracle(api, orc, dt, latname, grid, norb):
    from qmcpack_b200.workload import random_table
    lat = LATTICES[latname]
    G = np.linalg.inv(lat)
    coefs = random_table(grid, norb, dt, seed=3)
    spo = api.SplineSPOSet(coefs, norb, G)
    r = positions(lat, 37)
    psi, dpsi, d2psi = spo.mw_evaluateVGL(r)
    opsi, odpsi, od2psi = orc.r2r_vgl(coefs, G, norb, r)
    tol = TOL[np.dtype(dt)]
    assert rel_err(psi, opsi) < tol
    assert rel_err(dpsi, odpsi) < tol
    assert rel_err(d2psi, od2psi) < tol
    v = spo.mw_evaluateValue(r)
    assert rel_err(v, orc.r2r_value(coefs, G, norb, r)) < tol


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("norb", [24, 192, 384, 500])
def test_r2r_vgl_ratio_grads_matches_oracle(api, orc, dt, norb):
    from qmcpack_b200.workload import random_table
    lat = LATTICES["general"]
    G = np.linalg.inv(lat)
    coefs = random_table((6, 7, 5), norb, dt, seed=5)
    spo = api.SplineSPOSet(coefs, norb, G)
    nw = 21
    r = positions(lat, nw, seed=1)
    invrow = np.random.default_rng(2).normal(size=(nw, norb + 3)).astype(dt)  # ld_inv > n on purpose
    phi, ratios, grads = spo.mw_evaluateVGLandDetRatioGrads(r, invrow)
    ophi, oratios, ograds = orc.r2r_vgl_ratio_grads(coefs, G, norb, r, invrow)
    tol = TOL[np.dtype(dt)]
    assert rel_err(phi, ophi) < tol
    # the dot products cancel: compare against the size of the summed terms
    scale = np.abs(invrow[:, :norb]).max() * np.abs(ophi[0]).max() * np.sqrt(norb)
    assert np.abs(ratios - oratios).max() / scale < tol
    gscale = np.abs(ograds).max()
    assert np.abs(grads - ograds).max() / gscale < (2e-3 if dt == np.float32 else 1e-9)


@pytest.mark.parametrize("kind", ["R2R", "C2C"])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_offload_contract_device_rows_host_results(api, orc, dt, kind):
    """SPOSet::mw_evaluateVGLandDetRatioGrads as DiracDeterminantBatched::mw_ratioGrad drives an offload SPOSet
    (DiracDeterminantBatched.cpp:334-346): host positions, DEVICE inverse rows, phi_vgl_v current on the device on
    return, ratios / gradients on the host -- the same numbers as the host-buffer entry, bit for bit."""
    import torch
    from qmcpack_b200.workload import random_table
    lat = LATTICES["general"]
    G = np.linalg.inv(lat)
    norb, nw, ld = 130, 9, 136
    cplx = kind == "C2C"
    coefs = random_table((6, 7, 5), norb * (2 if cplx else 1), dt, seed=7)
    kc = np.random.default_rng(3).normal(size=(norb, 3)) * 0.3 if cplx else None
    spo = api.SplineSPOSet(coefs, norb, G, kind=api.C2C if cplx else api.R2R, kcart=kc)
    r = positions(lat, nw, seed=11)
    rng = np.random.default_rng(2)
    invrow = rng.normal(size=(nw, ld)).astype(dt)
    if cplx:
        invrow = (invrow + 1j * rng.normal(size=(nw, ld))).astype(spo.vt)
    phi, ratios, grads = spo.mw_evaluateVGLandDetRatioGrads(r, invrow)
    real_view = invrow.view(dt) if cplx else invrow
    inv_dev = torch.from_numpy(np.ascontiguousarray(real_view)).cuda()
    phi_dev = torch.zeros(phi.view(dt).shape if cplx else phi.shape, dtype=inv_dev.dtype, device="cuda")
    torch.cuda.synchronize()
    r2, g2 = spo.mw_evaluateVGLandDetRatioGrads_offload(r, inv_dev.data_ptr(), ld, phi_dev.data_ptr())
    assert np.array_equal(r2, ratios) and np.array_equal(g2, grads)
    got = phi_dev.cpu().numpy()
    assert np.array_equal(got.view(spo.vt).reshape(phi.shape) if cplx else got, phi)
    # and without the optional phi_vgl_v
    r3, g3 = spo.mw_evaluateVGLandDetRatioGrads_offload(r, inv_dev.data_ptr(), ld)
    assert np.array_equal(r3, ratios) and np.array_equal(g3, grads)


def test_r2r_ratio_is_deterministic(api):
    """fixed-order reduction: two evaluations give bit-identical ratios/gradients"""
    from qmcpack_b200.workload import random_table
    lat = LATTICES["cubic"]
    coefs = random_table((8, 8, 8), 384, np.float32, seed=9)
    spo = api.SplineSPOSet(coefs, 384, np.linalg.inv(lat))
    r = positions(lat, 64, seed=4)
    invrow = np.random.default_rng(3).normal(size=(64, 384)).astype(np.float32)
    a = spo.mw_evaluateVGLandDetRatioGrads(r, invrow)
    b = spo.mw_evaluateVGLandDetRatioGrads(r, invrow)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_r2r_halfG_sign(api, orc, dt):
    """real orbitals at half-G twists flip sign with the image parity (SplineR2R.h:156-170)"""
    from qmcpack_b200.workload import random_table
    lat = LATTICES["cubic"]
    G = np.linalg.inv(lat)
    coefs = random_table((6, 6, 6), 16, dt, seed=7)
    halfG = [1, 0, 1]
    spo = api.SplineSPOSet(coefs, 16, G, halfG=halfG)
    r = positions(lat, 30, seed=6)
    psi, dpsi, d2psi = spo.mw_evaluateVGL(r)
    opsi, odpsi, od2psi = orc.r2r_vgl(coefs, G, 16, r, halfG=halfG)
    tol = TOL[np.dtype(dt)]
    assert rel_err(psi, opsi) < tol and rel_err(dpsi, odpsi) < tol and rel_err(d2psi, od2psi) < tol


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_det_ratios_virtual_positions(api, orc, dt):
    """SPOSet::mw_evaluateDetRatios: V-only evaluation at quadrature points dotted with the reference walker's row"""
    from qmcpack_b200.workload import random_table
    lat = LATTICES["general"]
    G = np.linalg.inv(lat)
    norb, nw, nvp = 200, 5, 60
    coefs = random_table((6, 5, 7), norb, dt, seed=11)
    spo = api.SplineSPOSet(coefs, norb, G)
    rng = np.random.default_rng(8)
    r_vp = (rng.random((nvp, 3)) * 2 - 0.5) @ lat
    ref = rng.integers(0, nw, nvp)
    invrow = rng.normal(size=(nw, norb)).astype(dt)
    ratios = spo.mw_evaluateDetRatios(r_vp, ref, invrow)
    psi = orc.r2r_value(coefs, G, norb, r_vp)
    expect = np.einsum("ij,ij->i", psi.astype(np.float64), invrow[ref].astype(np.float64))
    scale = np.abs(invrow).max() * np.abs(psi).max() * np.sqrt(norb)
    assert np.abs(ratios - expect).max() / scale < TOL[np.dtype(dt)]


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_c2c_det_ratios_virtual_positions(api, orc, dt):
    """SPOSet::mw_evaluateDetRatios with complex orbitals (SplineC2COMPTarget.cpp:230-352): twisted V-only evaluation at
    the quadrature points, plain (unconjugated) complex dot with the reference walker's inverse row"""
    from qmcpack_b200.workload import random_table
    from qmcpack_b200.api import C2C
    lat = LATTICES["general"]
    G = np.linalg.inv(lat)
    norb, nw, nvp = 150, 4, 50
    coefs = random_table((6, 5, 7), 2 * norb, dt, seed=21)
    rng = np.random.default_rng(22)
    kcart = rng.normal(size=(norb, 3)) * 0.4
    spo = api.SplineSPOSet(coefs, norb, G, kind=C2C, kcart=kcart)
    r_vp = (rng.random((nvp, 3)) * 2 - 0.5) @ lat
    ref = rng.integers(0, nw, nvp)
    cdt = np.complex64 if dt == np.float32 else np.complex128
    invrow = (rng.normal(size=(nw, norb)) + 1j * rng.normal(size=(nw, norb))).astype(cdt)
    ratios = spo.mw_evaluateDetRatios(r_vp, ref, invrow)
    psi = orc.c2c_vgl(coefs, G, kcart, norb, r_vp, value_only=True)[0]
    expect = np.einsum("ij,ij->i", psi.astype(np.complex128), invrow[ref].astype(np.complex128))
    scale = np.abs(invrow).max() * np.abs(psi).max() * np.sqrt(norb)
    tol = TOL[np.dtype(dt)] * (4 if dt == np.float32 else 1)
    assert np.abs(ratios - expect).max() / scale < tol


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("norb", [12, 130, 384])
def test_c2c_matches_oracle(api, orc, dt, norb):
    """complex orbitals with a twist (SplineC2C.cpp:200-277 / ApplyPhaseC2C.hpp)"""
    from qmcpack_b200.workload import random_table
    from qmcpack_b200.api import C2C
    lat = LATTICES["general"]
    G = np.linalg.inv(lat)
    coefs = random_table((5, 6, 5), 2 * norb, dt, seed=13)
    rng = np.random.default_rng(14)
    kcart = rng.normal(size=(norb, 3)) * 0.4
    spo = api.SplineSPOSet(coefs, norb, G, kind=C2C, kcart=kcart)
    nw = 17
    r = positions(lat, nw, seed=15)
    psi, dpsi, d2psi = spo.mw_evaluateVGL(r)
    opsi, odpsi, od2psi = orc.c2c_vgl(coefs, G, kcart, norb, r)
    tol = TOL[np.dtype(dt)] * (4 if dt == np.float32 else 1)  # sincos of an O(10) phase in float
    assert rel_err(psi, opsi) < tol
    assert rel_err(dpsi, odpsi) < tol
    assert rel_err(d2psi, od2psi) < tol
    v = spo.mw_evaluateValue(r)
    assert rel_err(v, orc.c2c_vgl(coefs, G, kcart, norb, r, value_only=True)[0]) < tol
    cdt = np.complex64 if dt == np.float32 else np.complex128
    invrow = (rng.normal(size=(nw, norb)) + 1j * rng.normal(size=(nw, norb))).astype(cdt)
    phi, ratios, grads = spo.mw_evaluateVGLandDetRatioGrads(r, invrow)
    ophi, oratios, ograds = orc.c2c_vgl_ratio_grads(coefs, G, kcart, norb, r, invrow)
    assert rel_err(phi, ophi) < tol
    scale = np.abs(invrow).max() * np.abs(ophi[0]).max() * np.sqrt(norb)
    assert np.abs(ratios - oratios).max() / scale < tol


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_gpu_against_reference_golden_sine_grid(api, orc, dt):
    """the reference's own golden numbers (test_multi_spline.cpp:231-314) through the CUDA path"""
    import test_oracle_golden as tg
    coefs = tg.multi_table(orc.create_periodic_coefs(tg.sine_grid()), dt)
    spo = api.SplineSPOSet(coefs, 1, np.eye(3))
    psi, dpsi, d2psi = spo.mw_evaluateVGL([[0.1, 0.2, 0.3], [0.0, 0.0, 0.0]])
    tol = dict(rel=2e-5, abs=2e-4) if dt == np.float32 else dict(rel=1e-8, abs=1e-8)
    assert psi[0, 0] == pytest.approx(-0.9476393279, **tol)
    assert dpsi[0, 0] == pytest.approx([5.111042137, 5.989106342, 1.952244379], **tol)
    assert d2psi[0, 0] == pytest.approx(147.1127789, **tol)
    assert dpsi[1, 0] == pytest.approx([6.178320809, -7.402942564, -6.178320809], **tol)


def test_empty_batch_is_a_noop(api):
    from qmcpack_b200.workload import random_table
    coefs = random_table((4, 4, 4), 8, np.float32, seed=1)
    spo = api.SplineSPOSet(coefs, 8, np.eye(3))
    psi, dpsi, d2psi = spo.mw_evaluateVGL(np.zeros((0, 3)))
    assert psi.shape == (0, 8)


def test_bad_table_is_rejected(api):
    with pytest.raises(RuntimeError, match="npad"):
        api.SplineSPOSet(np.zeros((4, 4, 4, 10), np.float32), 4, np.eye(3))  # npad not a multiple of 64 bytes
