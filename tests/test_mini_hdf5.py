"""scripts/mini_hdf5.py (the HDF5 reader behind the diamondC fixture) against the reference's own ES-HDF file and the
committed fixture.  Needs /root/reference, so it only runs in the build container (skipped on the GPU box)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "scripts"))
SRC = "/root/reference/tests/solids/diamondC_1x1x1_pp/pwscf.pwscf.h5"
pytestmark = pytest.mark.skipif(not os.path.exists(SRC), reason="reference tree not present")


def test_groups_and_scalars():
    from mini_hdf5 import H5File
    f = H5File(SRC)
    assert sorted(f.listdir("/")) == ["application", "atoms", "electrons", "format", "supercell", "version"]
    assert f.read("/format")[0] == b"ES-HDF"
    assert list(f.read("/version")) == [2, 0, 0]
    assert list(f.read("/electrons/number_of_electrons")) == [4, 4]
    assert f.exists("/electrons/kpoint_0/spin_0/state_7/psi_g") and not f.exists("/electrons/psi_r_mesh")
    with pytest.raises(KeyError):
        f.read("/electrons/kpoint_0/spin_0/state_8/psi_g")


def test_fixture_matches_the_file():
    from mini_hdf5 import H5File
    f = H5File(SRC)
    d = np.load(os.path.join(HERE, "golden", "diamondC_1x1x1_eshdf.npz"))
    assert np.array_equal(d["gvectors"], f.read("/electrons/kpoint_0/gvectors"))
    assert np.array_equal(d["eigenvalues"], f.read("/electrons/kpoint_0/spin_0/eigenvalues"))
    assert np.array_equal(d["primitive_vectors"], f.read("/supercell/primitive_vectors"))
    for s in (0, 3, 7):
        c = f.read("/electrons/kpoint_0/spin_0/state_%d/psi_g" % s)
        assert np.array_equal(d["psi_g"][s], c[:, 0] + 1j * c[:, 1])
    # plane-wave coefficients of a normalised orbital
    assert abs(np.sum(np.abs(d["psi_g"][0])**2) - 1.0) < 1e-6


def test_second_reference_file_parses():
    """the 2x1x1 cell of the same test set: a different file, same reader"""
    from mini_hdf5 import H5File
    p = "/root/reference/tests/solids/diamondC_2x1x1_pp/pwscf.pwscf.h5"
    if not os.path.exists(p):
        pytest.skip("file not present")
    f = H5File(p)
    g = f.read("/electrons/kpoint_0/gvectors")
    assert g.ndim == 2 and g.shape[1] == 3
    nst = int(f.read("/electrons/kpoint_0/spin_0/number_of_states").ravel()[0])
    assert f.read("/electrons/kpoint_0/spin_0/state_%d/psi_g" % (nst - 1)).shape == (g.shape[0], 2)
