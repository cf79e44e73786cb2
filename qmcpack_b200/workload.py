"""Synthetic NiO-supercell workloads (BASELINE.json:configs).

Random-coefficient periodic spline tables of the named NiO shapes, the NiO J1/J2 B-spline functors of the
reference's performance input (tests/performance/NiO/sample/dmc-a64-e768-cpu/NiO-fcc-S16-dmc.xml.in:131-146)
and uniform random initial electron positions.  The precedent for random-coefficient tables is the reference's
own mini-app (src/Sandbox/einspline_spo.hpp:140-185).  Pure numpy: no device code here.
"""
import numpy as np

# J2 / J1 parameters of the NiO a64 benchmark input (reference xml cited above)
J2_UU = [0.28622356, 0.1947736865, 0.1319544873, 0.08893394669, 0.05695575776, 0.03565958405, 0.0220695026,
         0.01296086466, 0.006601006996, 0.00278714433]
J2_UD = [0.3689309537, 0.2226722029, 0.1484296802, 0.09617039126, 0.0591878654, 0.03660855878, 0.02262411664,
         0.01322279598, 0.006736329049, 0.002871931038]
J2_RCUT = 5.5727792532
J1_O = [-0.2249112633, -0.1847494689, -0.115481408, -0.04000122947, 0.01731711068, 0.05360131926, 0.05983040879,
        0.03955999983, 0.0173998007, 0.005162164083]
J1_NI = [-1.64485534, -1.470658909, -1.078893976, -0.6878964509, -0.3907004509, -0.1962103494, -0.08512755539,
         -0.02752356864, -0.00401798318, 0.0007665934444]
J1_RCUT = 4.8261684030
L_A64 = 15.7622  # cubic cell edge (bohr) of the a64 input

# name -> (N electrons, grid M, delay rank, dtype)
CONFIGS = {
    "NiO-a32": dict(N=384, M=48, k=32, dtype=np.float32),
    "NiO-a64": dict(N=768, M=60, k=32, dtype=np.float32),
    "NiO-a128": dict(N=1536, M=76, k=64, dtype=np.float64, complex_orbitals=True),  # full precision, SplineC2C
    "NiO-a256": dict(N=3072, M=96, k=32, dtype=np.float32),
}


def aligned_size(dtype, n):
    nd = 64 // np.dtype(dtype).itemsize
    return ((n + nd - 1) // nd) * nd


def aligned_zeros(shape, dtype, align=64):
    """numpy zeros whose data pointer is `align`-byte aligned (the reference CPU kernels use aligned SIMD loads)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    buf = np.zeros(n * dtype.itemsize + align, np.uint8)
    off = (-buf.ctypes.data) % align
    return buf[off:off + n * dtype.itemsize].view(dtype).reshape(shape)


def random_table(M, n_spl, dtype, seed, scale=0.5):
    """[M+3]^3 x npad coefficients, i.i.d. uniform(-scale, scale), periodic wrap C[M+i] = C[i] (i = 0..2) on every axis."""
    if np.isscalar(M):
        M = (M, M, M)
    npad = aligned_size(dtype, n_spl)
    rng = np.random.default_rng(seed)
    c = aligned_zeros((M[0] + 3, M[1] + 3, M[2] + 3, npad), dtype)
    core = rng.random((M[0], M[1], M[2], n_spl), dtype=np.float32 if np.dtype(dtype) == np.float32 else np.float64)
    core -= 0.5
    core *= 2 * scale
    c[:M[0], :M[1], :M[2], :n_spl] = core
    c[M[0]:, :, :, :] = c[:3, :, :, :]
    c[:, M[1]:, :, :] = c[:, :3, :, :]
    c[:, :, M[2]:, :] = c[:, :, :3, :]
    return c


def pw_table(M, n_spl, dtype, seed):
    """Smooth synthetic orbitals: a random ORTHOGONAL mixture (seeded) of the n_spl lowest plane-wave functions
    {1, cos(G.r), sin(G.r)} of the cell, sampled on the coefficient grid and used directly as B-spline coefficients.
    Same shape, strides and access pattern as any other table -- the gather kernels never look at the values -- but the
    Slater matrices are as well conditioned as a free-electron determinant's, so mixed-precision VMC stays meaningful
    (i.i.d. coefficients at 0.3 bohr spacing give a white-noise wavefunction whose inverse loses all digits within one
    sweep even in FP64)."""
    if np.isscalar(M):
        M = (M, M, M)
    npad = aligned_size(dtype, n_spl)
    rng = np.random.default_rng(seed)
    # integer reciprocal vectors, one of each +-G pair, sorted by |G|^2
    gmax = 1
    while (2 * gmax + 1) ** 3 < 2 * n_spl + 8:
        gmax += 1
    g = np.array([(i, j, k) for i in range(-gmax, gmax + 1) for j in range(-gmax, gmax + 1)
                  for k in range(-gmax, gmax + 1)])
    keep = [(i, j, k) for (i, j, k) in g if (i, j, k) > (0, 0, 0)]
    keep.sort(key=lambda t: (t[0] ** 2 + t[1] ** 2 + t[2] ** 2, t))
    funcs = [((0, 0, 0), 0)]
    for t in keep:
        funcs.append((t, 0))
        funcs.append((t, 1))
        if len(funcs) >= n_spl:
            break
    funcs = funcs[:n_spl]
    gv = np.array([f[0] for f in funcs], np.float64)
    is_sin = np.array([f[1] for f in funcs], bool)
    Q, _ = np.linalg.qr(rng.normal(size=(n_spl, n_spl)))
    Q = (Q * np.sqrt(2.0)).astype(np.float32 if np.dtype(dtype) == np.float32 else np.float64)
    c = aligned_zeros((M[0] + 3, M[1] + 3, M[2] + 3, npad), dtype)
    ys, zs = np.meshgrid(np.arange(M[1] + 3) / M[1], np.arange(M[2] + 3) / M[2], indexing="ij")
    pyz = 2 * np.pi * (ys.reshape(-1, 1) * gv[None, :, 1] + zs.reshape(-1, 1) * gv[None, :, 2])
    for ix in range(M[0]):
        ph = pyz + 2 * np.pi * (ix / M[0]) * gv[None, :, 0]
        basis = np.where(is_sin[None, :], np.sin(ph), np.cos(ph)).astype(Q.dtype)
        c[ix, :, :, :n_spl] = (basis @ Q).reshape(M[1] + 3, M[2] + 3, n_spl)
    c[M[0]:, :, :, :] = c[:3, :, :, :]
    return c


def pw_table_complex(M, n_orb, dtype, seed):
    """Complex analogue of pw_table for SplineC2C: orbital j = sum_G Q[G, j] exp(i G.r) with Q a seeded random UNITARY
    matrix over the n_orb lowest reciprocal vectors; real and imaginary parts are stored as the component pair
    (2j, 2j+1) of a [M+3]^3 x npad(2 n_orb) real table (SplineC2C.h: "the internal storage is real type arrays")."""
    if np.isscalar(M):
        M = (M, M, M)
    npad = aligned_size(dtype, 2 * n_orb)
    rng = np.random.default_rng(seed)
    gmax = 1
    while (2 * gmax + 1) ** 3 < n_orb + 8:
        gmax += 1
    g = [(i, j, k) for i in range(-gmax, gmax + 1) for j in range(-gmax, gmax + 1) for k in range(-gmax, gmax + 1)]
    g.sort(key=lambda t: (t[0] ** 2 + t[1] ** 2 + t[2] ** 2, t))
    gv = np.array(g[:n_orb], np.float64)
    Q, _ = np.linalg.qr(rng.normal(size=(n_orb, n_orb)) + 1j * rng.normal(size=(n_orb, n_orb)))
    c = aligned_zeros((M[0] + 3, M[1] + 3, M[2] + 3, npad), dtype)
    ys, zs = np.meshgrid(np.arange(M[1] + 3) / M[1], np.arange(M[2] + 3) / M[2], indexing="ij")
    pyz = 2 * np.pi * (ys.reshape(-1, 1) * gv[None, :, 1] + zs.reshape(-1, 1) * gv[None, :, 2])
    for ix in range(M[0]):
        ph = pyz + 2 * np.pi * (ix / M[0]) * gv[None, :, 0]
        u = (np.exp(1j * ph) @ Q).reshape(M[1] + 3, M[2] + 3, n_orb)
        c[ix, :, :, 0:2 * n_orb:2] = u.real
        c[ix, :, :, 1:2 * n_orb:2] = u.imag
    c[M[0]:, :, :, :] = c[:3, :, :, :]
    return c


def make_system(N=768, M=60, dtype=np.float32, L=None, seed=20240, with_j1=True, with_j2=True, lattice=None,
                same_table=False, table="pw", complex_orbitals=False, twist=(0.25, 0.25, 0.25)):
    """A synthetic NiO-like system: cubic cell scaled so the electron density matches a64, two spin tables.
    table = "pw" (random orthogonal plane-wave mixtures, default) or "iid" (i.i.d. uniform coefficients).
    complex_orbitals: SplineC2C tables (2n real components per spin) + one twist vector k = 2 pi G.twist for every
    orbital (system["kpts"], Cartesian) -- the NiO-a128 class of BASELINE.json."""
    n_up = N // 2
    n_dn = N - n_up
    if L is None:
        L = L_A64 * (N / 768.0) ** (1.0 / 3.0)
    lat = np.asarray(lattice, np.float64).reshape(3, 3) if lattice is not None else np.eye(3) * L
    if complex_orbitals:
        mk = pw_table_complex if table == "pw" else (lambda M_, n_, dt_, sd_: random_table(M_, 2 * n_, dt_, sd_))
    else:
        mk = pw_table if table == "pw" else random_table
    t_up = mk(M, n_up, dtype, seed)
    t_dn = t_up if same_table else mk(M, n_dn, dtype, seed + 1)
    s = dict(n_up=n_up, n_dn=n_dn, lattice=lat, coefs=[t_up, t_dn], grid=(M, M, M) if np.isscalar(M) else tuple(M))
    if complex_orbitals:
        kc = 2 * np.pi * (np.linalg.inv(lat) @ np.asarray(twist, np.float64))
        s["kpts"] = [np.tile(kc, (n_up, 1)), np.tile(kc, (n_dn, 1))]
    if with_j2:
        s["j2"] = dict(uu=J2_UU, ud=J2_UD, rcut=min(J2_RCUT, 0.4999 * L_wigner_seitz(lat)))
    if with_j1:
        nions = max(2, N // 12)
        rng = np.random.default_rng(seed + 17)
        frac = rng.random((nions, 3))
        s["j1"] = dict(ion_pos=frac @ lat, ion_grp=(np.arange(nions) % 2).astype(np.int32),
                       params=np.array([J1_O, J1_NI]), rcut=[min(J1_RCUT, 0.4999 * L_wigner_seitz(lat))] * 2)
    return s


def L_wigner_seitz(lat):
    """twice the inscribed-sphere radius of the cell (safe cutoff diameter for minimum-image functors)."""
    a = np.asarray(lat, np.float64)
    vol = abs(np.linalg.det(a))
    h = []
    for i in range(3):
        c = np.cross(a[(i + 1) % 3], a[(i + 2) % 3])
        h.append(vol / np.linalg.norm(c))
    return min(h)


def initial_positions(system, nw, seed=7):
    """uniform in the cell, default_rng(seed + walker) per walker."""
    N = system["n_up"] + system["n_dn"]
    lat = np.asarray(system["lattice"], np.float64)
    R = np.zeros((nw, N, 3))
    for iw in range(nw):
        R[iw] = np.random.default_rng(seed + iw).random((N, 3)) @ lat
    return R
