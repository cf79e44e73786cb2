"""Generates tests/golden/diamondC_1x1x1_eshdf.npz from the reference's own DFT orbital file
(/root/reference/tests/solids/diamondC_1x1x1_pp/pwscf.pwscf.h5, ES-HDF 2.0.0 written by pw2qmcpack) -- BASELINE.json
configs[0].  The reference tree does not travel to the GPU box and this image has no libhdf5/h5py, so the file is read
with scripts/mini_hdf5.py here and the plane-wave data the EinsplineSetBuilder consumes is committed as a fixture:
lattice, G-vectors, twist, eigenvalues and the psi_g coefficients of the 8 states of spin 0 (0.5 MB).
tests/test_diamondC_golden.py rebuilds the spline table from it (FFT -> phase fix -> periodic B-spline solve, following
BsplineFactory/OneSplineOrbData.cpp + einspline_helper.hpp) and checks the oracle and the CUDA kernels against the literals
of QMCWaveFunctions/tests/test_einset_diamondC.cpp.

  python scripts/gen_diamondC_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from mini_hdf5 import H5File

SRC = "/root/reference/tests/solids/diamondC_1x1x1_pp/pwscf.pwscf.h5"
f = H5File(SRC)
assert f.read("/format")[0] == b"ES-HDF"
nstates = int(f.read("/electrons/kpoint_0/spin_0/number_of_states").ravel()[0])
psi_g = np.stack([f.read("/electrons/kpoint_0/spin_0/state_%d/psi_g" % s) for s in range(nstates)])
psi_g = psi_g[..., 0] + 1j * psi_g[..., 1]
out = os.path.join(ROOT, "tests", "golden", "diamondC_1x1x1_eshdf.npz")
np.savez_compressed(out,
                    primitive_vectors=f.read("/supercell/primitive_vectors"),
                    gvectors=f.read("/electrons/kpoint_0/gvectors").astype(np.int32),
                    reduced_k=f.read("/electrons/kpoint_0/reduced_k"),
                    eigenvalues=f.read("/electrons/kpoint_0/spin_0/eigenvalues"),
                    psi_g=psi_g.astype(np.complex128),
                    version=f.read("/version"))
print(out, os.path.getsize(out), "bytes;", nstates, "states,", psi_g.shape[1], "G-vectors")


# ---- the 2x1x1 tiling of the same test set: two primitive-cell twists, (0,0,0) and (1/2,0,0); only the five lowest bands
# over both twists (what test_einset_diamondC.cpp:374-376 asks for) are kept, with their (twist, band) labels
SRC2 = "/root/reference/tests/solids/diamondC_2x1x1_pp/pwscf.pwscf.h5"
f2 = H5File(SRC2)
nk = int(f2.read("/electrons/number_of_kpoints").ravel()[0])
twists = np.stack([f2.read("/electrons/kpoint_%d/reduced_k" % k) for k in range(nk)])
eig = np.stack([f2.read("/electrons/kpoint_%d/spin_0/eigenvalues" % k) for k in range(nk)])
gv2 = f2.read("/electrons/kpoint_0/gvectors").astype(np.int32)
assert all(np.array_equal(gv2, f2.read("/electrons/kpoint_%d/gvectors" % k)) for k in range(nk))
order = sorted((round(float(eig[k][b]) / 1e-6) * 1e-6, k, b) for k in range(nk) for b in range(eig.shape[1]))[:5]
labels = np.array([(k, b) for _, k, b in order], np.int32)
pg = np.stack([f2.read("/electrons/kpoint_%d/spin_0/state_%d/psi_g" % (k, b)) for k, b in labels])
out2 = os.path.join(ROOT, "tests", "golden", "diamondC_2x1x1_eshdf.npz")
np.savez_compressed(out2, primitive_vectors=f2.read("/supercell/primitive_vectors"), gvectors=gv2, reduced_k=twists,
                    eigenvalues=eig, band_labels=labels, psi_g=(pg[..., 0] + 1j * pg[..., 1]).astype(np.complex128))
print(out2, os.path.getsize(out2), "bytes; bands (twist, band):", labels.tolist())
