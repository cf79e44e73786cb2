"""Minimum-image distance rows of the oracle in general cells (DTD_BConds<T,3,PPPG+SOA_OFFSET>,
/root/reference/src/Particle/Lattice/ParticleBConds3DSoa.h:339-510) against a brute-force search over image cells.

The reference reduces the lattice basis first (find_reduced_basis, LatticeAnalyzer.h:213-272) and builds the inverse and
the 8 corner shifts from the reduced basis; with the raw rows of a skewed, non-reduced cell the floor + 8-corner search
misses nearest images.  The brute force is independent of that algorithm."""
import numpy as np
import pytest

# an already reduced general cell, and two equivalent NON-reduced descriptions of lattices (a1 -> a1 + a0, a2 -> a2 + a0 - a1)
LAT_REDUCED = np.array([[6.0, 0.4, 0.0], [0.3, 6.5, -0.2], [0.1, -0.3, 7.0]])
SHEAR_1 = np.array([[1, 0, 0], [1, 1, 0], [0, 0, 1]], float)
SHEAR_2 = np.array([[1, 0, 0], [2, 1, 0], [1, -1, 1]], float)
FCC = 3.37 * np.array([[1, 1, 0], [0, 1, 1], [1, 0, 1]], float)
# name -> (cell as given to the code, an equivalent compact basis of the SAME lattice for the brute-force search)
LATS = {"reduced": (LAT_REDUCED, LAT_REDUCED), "a1+a0": (SHEAR_1 @ LAT_REDUCED, LAT_REDUCED),
        "sheared_twice": (SHEAR_2 @ LAT_REDUCED, LAT_REDUCED), "fcc_primitive": (FCC, FCC)}


def brute_force(lat, pos, src, nimg=3):
    """nearest image of src - pos: wrap the displacement into the cell of the compact basis, then search (2 nimg + 1)^3 cells"""
    rng = np.arange(-nimg, nimg + 1)
    shifts = np.array([(i, j, k) for i in rng for j in rng for k in rng], float) @ lat
    f = (src - pos) @ np.linalg.inv(lat)
    d = ((f - np.round(f)) @ lat)[None, :] + shifts
    r = np.linalg.norm(d, axis=1)
    i = np.argmin(r)
    return r[i], d[i]


@pytest.mark.parametrize("name", list(LATS))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_general_cell_rows_are_nearest_images(orc, name, dtype):
    lat, lat_search = LATS[name]
    rng = np.random.default_rng(11)
    N = 40
    npad = orc.aligned_size(dtype, N)
    R = rng.random((N, 3)) @ lat
    # sources spread over neighbouring cells too (positions are not wrapped by the drivers)
    R += rng.integers(-1, 2, size=(N, 3)) @ lat
    rsoa = np.zeros((3, npad), dtype)
    rsoa[:, :N] = R.T
    tol = 1e-10 if dtype == np.float64 else 2e-4
    for trial in range(6):
        pos = rng.random(3) @ lat
        for flip_ind in (0, N // 2, N):
            row = orc.dist_row(lat, pos, rsoa, N, flip_ind, dtype)
            for j in range(N):
                r, d = brute_force(lat_search, pos.astype(dtype).astype(float), rsoa[:, j].astype(float))
                assert row[0, j] == pytest.approx(r, abs=tol), (name, j, row[0, j], r)
                # the displacement is the nearest-image vector unless two images tie (never within tol here)
                assert np.allclose(row[1:, j], d, atol=tol * 10)


def test_oracle_driver_keeps_the_given_cell_for_the_orbitals(orc):
    """the spline SPOs convert positions with the inverse of the cell AS GIVEN (CrystalLattice::toUnit); only the distance
    tables work in the reduced basis.  For a non-reduced cell the two inverses differ: log|det| of the oracle driver must
    equal the determinant of the orbital matrix evaluated directly."""
    from qmcpack_b200 import workload
    import oracle_lib
    lat = LATS["sheared_twice"][0]
    s = workload.make_system(N=24, M=8, dtype=np.float64, L=6.0, lattice=lat, with_j1=False, with_j2=False)
    nw = 2
    R = workload.initial_positions(s, nw)
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[1], tau=0.1, delay_rank=4)
    ov.set_positions(R)
    ov.recompute()
    lp = ov.evaluate_gl()[0]
    G = np.linalg.inv(lat)
    for iw in range(nw):
        want = sum(np.linalg.slogdet(orc.r2r_vgl(s["coefs"][sp], G, 12, R[iw, sp * 12:(sp + 1) * 12])[0])[1] for sp in (0, 1))
        assert lp[iw] == pytest.approx(want, rel=1e-12)
