"""Plane waves -> 3-D B-spline table, the SPOSet construction step of the reference restated with numpy for
the parity tests.  Test infrastructure only.  Follows, for one twist of an ES-HDF orbital file:
  * mesh size from the G-vectors and meshfactor: BsplineFactory/EinsplineSetBuilderESHDF.fft.cpp:399-435
  * band order (energy, then band index): QMCWaveFunctions/BandInfo.h:50-61
  * unpack4fftw + backward FFT: BsplineFactory/einspline_helper.hpp:33-61, OneSplineOrbData.cpp:24-31,69-75
  * phase fix: fix_phase_rotate_c2r (einspline_helper.hpp:69-127, real orbitals / SplineR2R) and compute_phase +
    fix_phase_rotate_c2c (:157-185,235-279, complex orbitals / SplineC2C)
  * periodic coefficient solve in double, stored in the table's precision: einspline::set + MultiBspline::set_spline
    (done by the oracle's restatement of einspline's solve_periodic_interp_1d, checked against the reference-compiled
    create_UBspline_3d_d in tests/test_oracle_golden.py)."""
import numpy as np


def mesh_size(gvecs, meshfactor=1.0):
    m = [int(np.ceil(4.0 * meshfactor * x)) for x in np.abs(gvecs).max(0)]
    assert max(m) <= 128, "above 128 the reference rounds up to 2^a 3^b 5^c sizes (not needed for the fixtures)"
    return [x + x % 2 for x in m]


def fft_box(cg, gvecs, mesh):
    """unpack4fftw + FFTW_BACKWARD (unnormalised)"""
    mesh = np.asarray(mesh)
    ub = (mesh - 1) // 2
    lb = ub - mesh + 1
    ok = np.all((gvecs <= ub) & (gvecs >= lb), axis=1)
    idx = (gvecs[ok] + mesh) % mesh
    box = np.zeros(tuple(mesh), np.complex128)
    box[idx[:, 0], idx[:, 1], idx[:, 2]] = cg[ok]
    return np.fft.ifftn(box) * box.size


def _phase(box):
    r_norm = (box.real**2).sum()
    i_norm = (box.imag**2).sum()
    ri_norm = (box.real * box.imag).sum()
    x = (r_norm - i_norm) / ri_norm
    y = 1.0 / np.sqrt(x * x + 4.0)
    phs = np.sqrt(0.5 - y)
    return phs, (np.sqrt(1.0 - phs * phs) if x < 0 else -np.sqrt(1.0 - phs * phs))


def _eikr(mesh, twist):
    u = np.meshgrid(*[np.arange(m) / m for m in mesh], indexing="ij")
    return np.exp(2j * np.pi * (u[0] * twist[0] + u[1] * twist[1] + u[2] * twist[2]))


def rotate_c2r(box, twist):
    box = box * _eikr(box.shape, twist)
    pr, pi = _phase(box)
    return pr * box.real - pi * box.imag


def rotate_c2c(box, twist):
    pr, pi = _phase(box * _eikr(box.shape, twist))
    return box * (pr + 1j * pi)


def build_table(orc, psi_g, gvecs, twist, eigenvalues, norb, dtype, complex_orbitals=False, meshfactor=1.0):
    """[Mx+3][My+3][Mz+3][Npad] table of the `norb` lowest bands; complex orbitals take two components each"""
    mesh = mesh_size(gvecs, meshfactor)
    order = sorted(range(len(eigenvalues)), key=lambda b: (round(float(eigenvalues[b]) / 1e-6) * 1e-6, b))[:norb]
    ncomp = 2 * norb if complex_orbitals else norb
    npad = orc.aligned_size(dtype, ncomp)
    coefs = np.zeros((mesh[0] + 3, mesh[1] + 3, mesh[2] + 3, npad), dtype)
    for o, b in enumerate(order):
        box = fft_box(psi_g[b], gvecs, mesh)
        if complex_orbitals:
            z = rotate_c2c(box, twist)
            coefs[..., 2 * o] = orc.create_periodic_coefs(z.real).astype(dtype)
            coefs[..., 2 * o + 1] = orc.create_periodic_coefs(z.imag).astype(dtype)
        else:
            coefs[..., o] = orc.create_periodic_coefs(rotate_c2r(box, twist)).astype(dtype)
    return coefs


def band_order(eigenvalues, norb):
    """(twist, band) of the `norb` lowest states over all twists: BandInfo::operator< (energy within 1e-6, then twist,
    then band index), EinsplineSetBuilderCommon.cpp OccupyBands"""
    eig = np.atleast_2d(eigenvalues)
    keys = sorted((round(float(eig[k][b]) / 1e-6) * 1e-6, k, b) for k in range(eig.shape[0]) for b in range(eig.shape[1]))
    return [(k, b) for _, k, b in keys[:norb]]


def k_cart(G, twist):
    """BsplineSet::kPoints = PrimLattice.k_cart(-twist): the orbital is u(r) exp(+i k.r) and SplineC2C multiplies by
    exp(-i kPoint.r) (SplineC2C.cpp:146-168); G = inverse primitive vectors (ru = r G)"""
    return -2.0 * np.pi * (np.asarray(G) @ np.asarray(twist, np.float64))


def build_table_c2c_twists(orc, psi_g, band_twist, twists, gvecs, dtype, meshfactor=1.0):
    """complex (SplineC2C) table of orbitals that come from several primitive-cell twists: psi_g[o] belongs to twist
    band_twist[o]; returns (table, kcart [norb][3] in reduced-to-Cartesian form still to be multiplied by G)"""
    mesh = mesh_size(gvecs, meshfactor)
    norb = len(psi_g)
    npad = orc.aligned_size(dtype, 2 * norb)
    coefs = np.zeros((mesh[0] + 3, mesh[1] + 3, mesh[2] + 3, npad), dtype)
    for o in range(norb):
        z = rotate_c2c(fft_box(psi_g[o], gvecs, mesh), twists[band_twist[o]])
        coefs[..., 2 * o] = orc.create_periodic_coefs(z.real).astype(dtype)
        coefs[..., 2 * o + 1] = orc.create_periodic_coefs(z.imag).astype(dtype)
    return coefs
