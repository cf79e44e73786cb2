"""BASELINE.json configs[0] (tests/solids/diamondC_1x1x1_pp) pinned at the SPOSet level: the reference's own DFT orbitals
(tests/golden/diamondC_1x1x1_eshdf.npz, extracted from pwscf.pwscf.h5 by scripts/gen_diamondC_golden.py) are turned into
the B-spline table exactly as EinsplineSetBuilder does (tests/eshdf_spline.py) and evaluated at the positions of
QMCWaveFunctions/tests/test_einset_diamondC.cpp:58-62 in the fcc primitive cell (a general, non-orthorhombic lattice).
Expected values are that test's literals: :98-113 (value, gradient, Laplacian), :128-136 (Hessian), :232-235 (batched
evaluate), :249-268 (mw_evaluateVGLandDetRatioGrads with inv_row = {0.1..0.5}, real build and QMC_COMPLEX build) and
:392-419 (the 2x1x1 tiling: SplineC2C orbitals from two primitive-cell twists, diamondC_2x1x1_eshdf.npz).
The reference checks with Catch's Approx (relative 100 * float epsilon ~ 1.2e-5 scaled by the value).
CPU tests run the oracle; the gpu tests run the CUDA kernels through the C ABI on the same table."""
import os

import numpy as np
import pytest

import eshdf_spline
import oracle_lib

HERE = os.path.dirname(os.path.abspath(__file__))
R_LAT = np.array([[3.37316115, 3.37316115, 0.0], [0.0, 3.37316115, 3.37316115], [3.37316115, 0.0, 3.37316115]])
POS = np.array([[0.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
REL = 2e-5  # Catch Approx default epsilon is 1.19e-5 relative; a little head room for the different FFT library


def approx(v, rel=REL, abs_=2e-5):
    return pytest.approx(v, rel=rel, abs=abs_)


@pytest.fixture(scope="module")
def orc():
    oracle_lib.build()
    return oracle_lib.port()


@pytest.fixture(scope="module")
def data():
    return np.load(os.path.join(HERE, "golden", "diamondC_1x1x1_eshdf.npz"))


@pytest.fixture(scope="module")
def table_real(orc, data):
    assert np.allclose(data["primitive_vectors"], R_LAT)
    return eshdf_spline.build_table(orc, data["psi_g"], data["gvectors"], data["reduced_k"], data["eigenvalues"], 8,
                                    np.float32)


@pytest.fixture(scope="module")
def table_cplx(orc, data):
    return eshdf_spline.build_table(orc, data["psi_g"], data["gvectors"], data["reduced_k"], data["eigenvalues"], 8,
                                    np.float32, complex_orbitals=True)


def check_vgl(psi, dpsi, d2psi):
    # test_einset_diamondC.cpp:98-113
    assert psi[0][0] == approx(-0.42546836868)
    assert psi[0][1] == approx(0.0)
    assert psi[1][0] == approx(-0.8886948824)
    assert psi[1][1] == approx(1.419412370359)
    assert dpsi[1][0][0] == approx(-0.0000183403)
    assert dpsi[1][0][1] == approx(0.1655139178)
    assert dpsi[1][0][2] == approx(-0.0000193077)
    assert dpsi[1][1][0] == approx(-1.3131694794)
    assert dpsi[1][1][1] == approx(-1.1174004078)
    assert dpsi[1][1][2] == approx(-0.8462534547)
    assert d2psi[1][0] == approx(1.3313053846, rel=2e-5 * 2)  # .epsilon(2e-5) in the reference
    assert d2psi[1][1] == approx(-4.712583065)


def check_ratio_grads_real(ratios, grads):
    # test_einset_diamondC.cpp:260-267
    assert ratios[0] == approx(-0.0425468457)
    assert grads[0] == approx([101.2666081556, 46.1671284048, 160.644753288])
    assert ratios[1] == approx(-0.5234490454)
    assert grads[1] == approx([1.8766445844, -0.513164153, 2.5277422458])


def check_ratio_grads_cplx(ratios, grads):
    # test_einset_diamondC.cpp:250-257 (QMC_COMPLEX); ComplexApprox default epsilon 100 * float epsilon
    def capprox(z):
        return pytest.approx(z, rel=1e-4, abs=2e-5)
    assert ratios[0] == capprox(complex(-0.0425468, 0.0425468))
    assert grads[0][0] == capprox(complex(99.0451, 2.22151))
    assert grads[0][1] == capprox(complex(52.8267, -6.65955))
    assert grads[0][2] == capprox(complex(156.207, 4.43802))
    assert ratios[1] == capprox(complex(-0.523449, 0.641483))
    assert grads[1][0] == capprox(complex(1.59725, 0.227989))
    assert grads[1][1] == capprox(complex(-0.543437, 0.0247015))
    assert grads[1][2] == capprox(complex(2.15262, 0.306102))


def test_mesh_and_bands(data):
    assert eshdf_spline.mesh_size(data["gvectors"]) == [40, 40, 40]
    assert data["psi_g"].shape == (8, 3695)


def test_oracle_vgl_matches_reference_literals(orc, table_real):
    G = np.linalg.inv(R_LAT)
    psi, dpsi, d2psi = orc.r2r_vgl(table_real, G, 8, POS)
    check_vgl(psi, dpsi, d2psi)
    # batched evaluate with the positions interchanged (:232-235)
    psi2, _, _ = orc.r2r_vgl(table_real, G, 8, POS[::-1])
    assert psi2[0][0] == approx(-0.8886948824) and psi2[1][0] == approx(-0.42546836868)


def test_oracle_hessian_matches_reference_literals(orc, table_real):
    """evaluateVGH at (0,1,0), orbital 1 (:128-136): Cartesian Hessian = G h G^T of the lattice-unit spline Hessian"""
    G = np.linalg.inv(R_LAT)
    ru = POS[1] @ G
    ru -= np.floor(ru)
    _, _, h = orc.spline_eval(table_real, 8, 2, ru)
    hm = np.array([[h[0][1], h[1][1], h[2][1]], [h[1][1], h[3][1], h[4][1]], [h[2][1], h[4][1], h[5][1]]], np.float64)
    hc = G @ hm @ G.T
    want = np.array([[-2.3160984034, 1.8089479397, 0.5608575749], [1.8089479397, -0.07996207476, 0.5237969314],
                     [0.5608575749, 0.5237969314, -2.316497764]])
    assert hc == approx(want, rel=2000 * 1.2e-7, abs_=3e-4)


def test_oracle_ratio_grads_real(orc, table_real):
    inv = np.tile(np.array([0.1, 0.2, 0.3, 0.4, 0.5], np.float32), (2, 1))
    _, ratios, grads = orc.r2r_vgl_ratio_grads(table_real, np.linalg.inv(R_LAT), 5, POS, inv)
    check_ratio_grads_real(ratios, grads)


def test_oracle_ratio_grads_complex(orc, table_cplx):
    inv = np.tile(np.array([0.1, 0.2, 0.3, 0.4, 0.5], np.complex64), (2, 1))
    _, ratios, grads = orc.c2c_vgl_ratio_grads(table_cplx, np.linalg.inv(R_LAT), np.zeros((5, 3)), 5, POS, inv)
    check_ratio_grads_cplx(ratios, grads)


def test_reference_compiled_kernels_on_the_real_table(orc, table_real):
    """the reference's own spline2 kernels (oracle/_ref, where built) agree with the restatement on the DFT table"""
    try:
        ref = oracle_lib.ref()
    except Exception:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    G = np.linalg.inv(R_LAT)
    for p in POS:
        ru = p @ G
        ru -= np.floor(ru)
        a = orc.spline_eval(table_real, 8, 2, ru)
        b = ref.spline_eval(table_real, 8, 2, ru)
        for x, y in zip(a, b):
            assert np.allclose(x, y, rtol=1e-6, atol=1e-6)


# ---------------------------------------------------------------- 2x1x1 tiling: orbitals from two primitive-cell twists
R_PRIM = R_LAT  # the spline lives on the primitive cell (BsplineSet::PrimLattice); the supercell only tiles it


@pytest.fixture(scope="module")
def data2():
    return np.load(os.path.join(HERE, "golden", "diamondC_2x1x1_eshdf.npz"))


@pytest.fixture(scope="module")
def table_2x1x1(orc, data2):
    labels = [tuple(x) for x in data2["band_labels"]]
    assert eshdf_spline.band_order(data2["eigenvalues"], 5) == labels == [(0, 0), (1, 0), (1, 1), (1, 2), (1, 3)]
    assert eshdf_spline.mesh_size(data2["gvectors"]) == [44, 40, 40]
    tw = [k for k, _ in labels]
    coefs = eshdf_spline.build_table_c2c_twists(orc, data2["psi_g"], tw, data2["reduced_k"], data2["gvectors"], np.float32)
    G = np.linalg.inv(R_PRIM)
    kcart = np.stack([eshdf_spline.k_cart(G, data2["reduced_k"][k]) for k in tw])
    return coefs, kcart


def check_2x1x1(psi, dpsi, d2psi):
    # test_einset_diamondC.cpp:392-419 (QMC_COMPLEX build: real and imaginary parts), electron 1 at (0, 1, 0)
    def c(re, im):
        return pytest.approx(complex(re, im), rel=2e-5, abs=2e-5)
    assert psi[1][0] == c(0.9008999467, 0.9008999467)
    assert psi[1][1] == c(1.2383049726, 1.2383049726)
    assert dpsi[1][0][0] == c(0.0025820041, 0.0025820041)
    assert dpsi[1][0][1] == c(-0.1880052537, -0.1880052537)
    assert dpsi[1][0][2] == c(-0.0025404284, -0.0025404284)
    assert dpsi[1][1][0] == c(0.1069662273, 0.1069453433)
    assert dpsi[1][1][1] == c(-0.4364597797, -0.43649593)
    assert dpsi[1][1][2] == c(-0.106951952, -0.1069145575)
    assert d2psi[1][0] == c(-1.3757134676, -1.3757134676)
    assert d2psi[1][1] == c(-2.4803137779, -2.4919104576)


POS5 = np.array([[0.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 1.1, 0.0], [0.0, 1.2, 0.0], [0.0, 1.3, 0.0]])


def test_oracle_two_twists_complex(orc, table_2x1x1):
    coefs, kcart = table_2x1x1
    psi, dpsi, d2psi = orc.c2c_vgl(coefs, np.linalg.inv(R_PRIM), kcart, 5, POS5)
    check_2x1x1(psi, dpsi, d2psi)


# ---------------------------------------------------------------- CUDA kernels through the C ABI
@pytest.fixture(scope="module")
def api():
    from qmcpack_b200 import api as a, build
    build.build()
    a.init(0)
    return a


@pytest.mark.gpu
def test_gpu_vgl_matches_reference_literals(api, orc, table_real):
    G = np.linalg.inv(R_LAT)
    spo = api.SplineSPOSet(table_real, 8, G)
    psi, dpsi, d2psi = spo.mw_evaluateVGL(POS)
    check_vgl(psi, dpsi, d2psi)
    opsi, odpsi, od2psi = orc.r2r_vgl(table_real, G, 8, POS)
    for a, b in ((psi, opsi), (dpsi, odpsi), (d2psi, od2psi)):
        assert np.abs(a - b).max() <= 1e-5 * np.abs(b).max()
    v = spo.mw_evaluateValue(POS[::-1])
    assert v[0][0] == approx(-0.8886948824) and v[1][0] == approx(-0.42546836868)


@pytest.mark.gpu
def test_gpu_ratio_grads_real(api, table_real):
    spo = api.SplineSPOSet(table_real, 8, np.linalg.inv(R_LAT))
    inv = np.zeros((2, 8), np.float32)
    inv[:, :5] = [0.1, 0.2, 0.3, 0.4, 0.5]  # the reference asks for 5 of the 8 orbitals; zero weights on the rest
    out = spo.mw_evaluateVGLandDetRatioGrads(POS, inv)
    ratios, grads = out[-2], out[-1]
    check_ratio_grads_real(ratios, grads)


@pytest.mark.gpu
def test_gpu_ratio_grads_complex(api, table_cplx):
    spo = api.SplineSPOSet(table_cplx, 8, np.linalg.inv(R_LAT), kind=api.C2C, kcart=np.zeros((8, 3)))
    inv = np.zeros((2, 8), np.complex64)
    inv[:, :5] = [0.1, 0.2, 0.3, 0.4, 0.5]
    out = spo.mw_evaluateVGLandDetRatioGrads(POS, inv)
    ratios, grads = out[-2], out[-1]
    check_ratio_grads_cplx(ratios, grads)


@pytest.mark.gpu
def test_gpu_two_twists_complex(api, orc, table_2x1x1):
    coefs, kcart = table_2x1x1
    G = np.linalg.inv(R_PRIM)
    spo = api.SplineSPOSet(coefs, 5, G, kind=api.C2C, kcart=kcart)
    psi, dpsi, d2psi = spo.mw_evaluateVGL(POS5)
    check_2x1x1(psi, dpsi, d2psi)
    opsi, odpsi, od2psi = orc.c2c_vgl(coefs, G, kcart, 5, POS5)
    for a, b in ((psi, opsi), (dpsi, odpsi), (d2psi, od2psi)):
        assert np.abs(a - b).max() <= 1e-5 * np.abs(b).max()
