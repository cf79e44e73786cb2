"""Reader for the B-spline coefficient dumps QMCPACK writes with `save_coefs="yes"` and restores on the next run
(BsplineFactory/SplineSetReader.cpp:130-180: datasets `class_name`, `sizeof`, and one 4-D table per block through
SplineUtils<ST>::write, spline2/SplineUtils.cpp:37-51; the table's HDF5 shape is
[x_grid.num + 3][y_grid.num + 3][z_grid.num + 3][z_stride], spline2/einspline_engine.hpp:44-52 -- exactly the
multi_UBspline_3d_{s,d} block `qmcb_spline_create` takes).  A file produced by a QMCPACK run of the same input therefore
gives this library the REAL orbitals of a benchmark without the plane-wave -> spline construction.

File names: `<root>.spin_<s>.tile_<t...>.tw_<i>.g<N>x<N>x<N>.h5` style per band group (BsplineReader::getSplineDumpFileName);
spin-up and spin-down groups are separate files.

No libhdf5 in this image: the file is parsed by qmcpack_b200/mini_hdf5.py (version-0 superblock subset, what libhdf5
writes by default for such a flat file)."""
import numpy as np

from .mini_hdf5 import H5File

_COMPLEX_CLASSES = ("SplineC2C", "SplineC2COMPTarget", "SplineC2R", "SplineC2ROMPTarget")


def _block_names(listing):
    """`spline_0`, then -- SplineUtils builds the name in an ostringstream it never clears -- `spline_0spline_1`, ...;
    plain `spline_<i>` is accepted as well."""
    names, i, acc = [], 0, ""
    while True:
        acc += "spline_%d" % i
        if acc in listing:
            names.append(acc)
        elif "spline_%d" % i in listing and i > 0:
            names.append("spline_%d" % i)
        else:
            break
        i += 1
    return names


def read_spline_dump(path):
    """-> dict(class_name, sizeof, dtype, blocks=[ndarray [Nx+3][Ny+3][Nz+3][z_stride]], grid=(Nx, Ny, Nz))"""
    f = H5File(path)
    listing = f.listdir("/")
    if "sizeof" not in listing or "spline_0" not in listing:
        raise ValueError(path + ": not a QMCPACK spline dump (datasets `sizeof` and `spline_0` expected)")
    sizeof = int(np.asarray(f.read("/sizeof")).reshape(-1)[0])
    cn = f.read("/class_name") if "class_name" in listing else np.array(b"")
    class_name = bytes(np.asarray(cn).reshape(-1)[0]).split(b"\0")[0].decode()
    blocks = []
    for name in _block_names(listing):
        a = f.read("/" + name)
        if a.ndim != 4 or a.dtype.kind != "f" or a.dtype.itemsize != sizeof:
            raise ValueError("%s:%s: expected a 4-D table of %d-byte reals, found %s %s" %
                             (path, name, sizeof, a.shape, a.dtype))
        blocks.append(np.ascontiguousarray(a.astype(a.dtype.newbyteorder("="), copy=False)))
    g = blocks[0].shape
    return dict(class_name=class_name, sizeof=sizeof, dtype=blocks[0].dtype, blocks=blocks,
                grid=(g[0] - 3, g[1] - 3, g[2] - 3), is_complex=class_name in _COMPLEX_CLASSES)


def sposet_from_dump(path, n_orb, G, halfG=None, kcart=None, block=0):
    """SplineSPOSet over the table of one dump file.  n_orb, the primitive-cell G, HalfG and the orbitals' k-points are
    not stored in the dump (the reference re-derives them from the input and the ES-HDF file, BsplineSet.h:38-252): the
    caller supplies them, as the in-tree adapter does (integration/SplineB200.h)."""
    from . import api
    d = read_spline_dump(path)
    coefs = d["blocks"][block]
    kind = api.C2C if d["is_complex"] else api.R2R
    need = (2 if kind == api.C2C else 1) * n_orb
    if coefs.shape[3] < need:
        raise ValueError("%s holds %d components per grid point, %d orbitals need %d" % (path, coefs.shape[3], n_orb, need))
    return api.SplineSPOSet(coefs, n_orb, G, kind=kind, halfG=halfG, kcart=kcart)
