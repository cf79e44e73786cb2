"""ctypes bindings for the CPU oracle (oracle/liboracle.so and, when built, oracle/_ref/libqmcref.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product package (qmcpack_b200) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

# the reference runs BLAS single-threaded inside each crowd thread (BlasThreadingEnv, DelayedUpdate.h:163-170): the
# OpenBLAS that oracle/_ref links must not spawn its own pool under the OpenMP crowd loop
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PORT_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libqmcref.so")

c_dp = C.POINTER(C.c_double)
c_fp = C.POINTER(C.c_float)
c_ip = C.POINTER(C.c_int)


def build(force=False):
    """Compile the oracle with oracle/Makefile (g++ only).  oracle/_ref is (re)built only where /root/reference exists."""
    if force or not os.path.exists(PORT_SO):
        subprocess.check_call(["make", "-C", ORACLE_DIR, PORT_SO], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src") and (force or not os.path.exists(REF_SO)):
        subprocess.call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


class VMCParams(C.Structure):
    _fields_ = [
        ("precision", C.c_int), ("n_up", C.c_int), ("n_dn", C.c_int),
        ("lattice", C.c_double * 9),
        ("coefs", C.c_void_p * 2), ("grid", C.c_int * 3), ("npad", C.c_int),
        ("n_j2", C.c_int), ("j2_uu", c_dp), ("j2_ud", c_dp), ("j2_rcut", C.c_double),
        ("nions", C.c_int), ("ion_pos", c_dp), ("ion_grp", c_ip), ("n_ion_groups", C.c_int),
        ("n_j1", C.c_int), ("j1_params", c_dp), ("j1_rcut", c_dp),
        ("nw", C.c_int), ("ncrowds", C.c_int), ("seeds", C.POINTER(C.c_uint32)),
        ("tau", C.c_double), ("use_drift", C.c_int), ("delay_rank", C.c_int), ("batched_engine", C.c_int),
        ("complex_orbitals", C.c_int), ("kpts", c_dp * 2), ("dmc", C.c_int),
        ("has_spline_lattice", C.c_int), ("spline_lattice", C.c_double * 9),
    ]


def azeros(shape, dtype, align=64):
    """64-byte aligned zeros: the reference kernels in oracle/_ref use aligned SIMD loads/stores."""
    dtype = np.dtype(dtype)
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    n = int(np.prod(shape))
    buf = np.zeros(n * dtype.itemsize + align, np.uint8)
    off = (-buf.ctypes.data) % align
    return buf[off:off + n * dtype.itemsize].view(dtype).reshape(shape)


def _np(a, dt):
    a = np.ascontiguousarray(a, dtype=dt)
    if a.ctypes.data % 64:
        b = azeros(a.shape, a.dtype)
        b[...] = a
        a = b
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Thin numpy-facing wrapper around one oracle shared library."""

    def __init__(self, path):
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        L.orc_last_error.restype = C.c_char_p
        L.orc_symtrace.restype = C.c_double
        L.orc_symtrace.argtypes = [C.c_double] * 6 + [c_dp]
        L.orc_aligned_size.restype = C.c_long
        L.orc_aligned_size.argtypes = [C.c_int, C.c_long]
        L.orc_vmc_create.restype = C.c_void_p
        L.orc_vmc_create.argtypes = [C.POINTER(VMCParams)]
        for name in ("orc_du_create_d", "orc_du_create_f", "orc_rng_create"):
            getattr(L, name).restype = C.c_void_p
        L.orc_rng_next.restype = C.c_double
        L.orc_rng_next.argtypes = [C.c_void_p]
        self.is_reference = bool(L.orc_is_reference_build())

    # -- helpers
    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.orc_last_error().decode())

    @staticmethod
    def suf(dtype):
        return {np.dtype(np.float32): "f", np.dtype(np.float64): "d", np.dtype(np.complex64): "c",
                np.dtype(np.complex128): "z"}[np.dtype(dtype)]

    def aligned_size(self, dtype, n):
        """getAlignedSize<T>(n): 64-byte rows (aligned_allocator.hpp:41-47)"""
        per = 64 // np.dtype(dtype).itemsize
        return ((int(n) + per - 1) // per) * per

    # -- golden-vector pieces
    def symtrace(self, h, gg):
        g = _np(gg, np.float64)
        return self.lib.orc_symtrace(*[C.c_double(x) for x in h], g.ctypes.data_as(c_dp))

    def prefactors(self, t, dtype=np.float64):
        a, da, d2a = (np.zeros(4, dtype) for _ in range(3))
        ct = C.c_float if self.suf(dtype) == "f" else C.c_double
        getattr(self.lib, "orc_prefactors_" + self.suf(dtype))(ct(t), _p(a), _p(da), _p(d2a))
        return a, da, d2a

    def create_periodic_coefs(self, data):
        data = _np(data, np.float64)
        M = np.array(data.shape, dtype=np.int32)
        coefs = np.zeros(tuple(int(m) + 3 for m in M), np.float64)
        self.lib.orc_create_periodic_coefs_3d(_p(M), _p(data), _p(coefs))
        return coefs

    def spline_eval(self, coefs, ns, which, pos):
        """coefs [Nx][Ny][Nz][npad]; pos unit coords; returns (v, g, h) with g/h shaped [3|6][npad] ([3][npad] twice for VGL)."""
        dt = coefs.dtype
        grid = np.array([s - 3 for s in coefs.shape[:3]], np.int32)
        npad = coefs.shape[3]
        coefs = _np(coefs, dt)
        v = azeros(npad, dt)
        g = azeros((3, npad), dt)
        h = azeros((6 if which == 2 else 3, npad), dt)
        ct = C.c_float if dt == np.float32 else C.c_double
        f = getattr(self.lib, "orc_spline_eval_" + self.suf(dt))
        self._chk(f(_p(coefs), _p(grid), C.c_int(ns), C.c_long(npad), C.c_int(which), ct(pos[0]), ct(pos[1]),
                    ct(pos[2]), _p(v), _p(g), _p(h)))
        return v, g, h

    # -- SPOSet level
    def r2r_vgl(self, coefs, G, norb, r, halfG=None):
        dt = coefs.dtype
        grid = np.array([s - 3 for s in coefs.shape[:3]], np.int32)
        npad = coefs.shape[3]
        r = _np(r, np.float64)
        nw = r.shape[0]
        G = _np(G, np.float64)
        hg = _np(halfG if halfG is not None else [0, 0, 0], np.int32)
        psi = np.zeros((nw, norb), dt)
        dpsi = np.zeros((nw, norb, 3), dt)
        d2psi = np.zeros((nw, norb), dt)
        f = getattr(self.lib, "orc_r2r_mw_evaluate_vgl_" + self.suf(dt))
        self._chk(f(_p(coefs), _p(grid), C.c_int(norb), C.c_long(npad), _p(G), _p(hg), C.c_int(norb), C.c_int(nw),
                    _p(r), _p(psi), _p(dpsi), _p(d2psi)))
        return psi, dpsi, d2psi

    def r2r_value(self, coefs, G, norb, r, halfG=None):
        dt = coefs.dtype
        grid = np.array([s - 3 for s in coefs.shape[:3]], np.int32)
        npad = coefs.shape[3]
        r = _np(r, np.float64)
        nw = r.shape[0]
        G = _np(G, np.float64)
        hg = _np(halfG if halfG is not None else [0, 0, 0], np.int32)
        psi = np.zeros((nw, norb), dt)
        f = getattr(self.lib, "orc_r2r_mw_evaluate_value_" + self.suf(dt))
        self._chk(f(_p(coefs), _p(grid), C.c_int(norb), C.c_long(npad), _p(G), _p(hg), C.c_int(norb), C.c_int(nw),
                    _p(r), _p(psi)))
        return psi

    def r2r_vgl_ratio_grads(self, coefs, G, norb, r, invrow, halfG=None):
        dt = coefs.dtype
        grid = np.array([s - 3 for s in coefs.shape[:3]], np.int32)
        npad = coefs.shape[3]
        r = _np(r, np.float64)
        nw = r.shape[0]
        G = _np(G, np.float64)
        hg = _np(halfG if halfG is not None else [0, 0, 0], np.int32)
        invrow = _np(invrow, dt)
        phi = np.zeros((5, nw, norb), dt)
        ratios = np.zeros(nw, dt)
        grads = np.zeros((nw, 3), dt)
        f = getattr(self.lib, "orc_r2r_mw_vgl_ratio_grads_" + self.suf(dt))
        self._chk(f(_p(coefs), _p(grid), C.c_int(norb), C.c_long(npad), _p(G), _p(hg), C.c_int(norb), C.c_int(nw),
                    _p(r), _p(invrow), C.c_long(invrow.shape[1]), _p(phi), _p(ratios), _p(grads)))
        return phi, ratios, grads

    def c2c_vgl(self, coefs, G, kcart, norb, r, value_only=False):
        dt = coefs.dtype
        cdt = np.complex64 if dt == np.float32 else np.complex128
        grid = np.array([s - 3 for s in coefs.shape[:3]], np.int32)
        npad = coefs.shape[3]
        r = _np(r, np.float64)
        nw = r.shape[0]
        G = _np(G, np.float64)
        k = _np(kcart, np.float64)
        psi = np.zeros((nw, norb), cdt)
        dpsi = np.zeros((nw, norb, 3), cdt)
        d2psi = np.zeros((nw, norb), cdt)
        f = getattr(self.lib, "orc_c2c_mw_evaluate_vgl_" + self.suf(dt))
        self._chk(f(_p(coefs), _p(grid), C.c_int(2 * norb), C.c_long(npad), _p(G), _p(k), C.c_int(norb), C.c_int(nw),
                    _p(r), _p(psi), _p(dpsi), _p(d2psi), C.c_int(1 if value_only else 0)))
        return (psi,) if value_only else (psi, dpsi, d2psi)

    def c2c_vgl_ratio_grads(self, coefs, G, kcart, norb, r, invrow):
        dt = coefs.dtype
        cdt = np.complex64 if dt == np.float32 else np.complex128
        grid = np.array([s - 3 for s in coefs.shape[:3]], np.int32)
        npad = coefs.shape[3]
        r = _np(r, np.float64)
        nw = r.shape[0]
        G = _np(G, np.float64)
        k = _np(kcart, np.float64)
        invrow = _np(invrow, cdt)
        phi = np.zeros((5, nw, norb), cdt)
        ratios = np.zeros(nw, cdt)
        grads = np.zeros((nw, 3), cdt)
        f = getattr(self.lib, "orc_c2c_mw_vgl_ratio_grads_" + self.suf(dt))
        self._chk(f(_p(coefs), _p(grid), C.c_int(2 * norb), C.c_long(npad), _p(G), _p(k), C.c_int(norb), C.c_int(nw),
                    _p(r), _p(invrow), C.c_long(invrow.shape[1]), _p(phi), _p(ratios), _p(grads)))
        return phi, ratios, grads

    # -- dense inverse
    def invert_transpose(self, a, lda=None):
        a = np.ascontiguousarray(a)
        dt = a.dtype
        n = a.shape[0]
        lda = lda or n
        inv = np.zeros((n, lda), dt)
        logdet = np.zeros(2, np.float64)
        f = getattr(self.lib, "orc_invert_transpose_" + self.suf(dt))
        self._chk(f(_p(a), C.c_int(n), C.c_int(a.shape[1]), _p(inv), C.c_int(lda), _p(logdet)))
        return inv, complex(logdet[0], logdet[1])

    # -- delayed update engine
    def du(self, n, k, dtype=np.float64):
        return DelayedUpdateHandle(self, n, k, dtype)

    # -- Jastrow / distances
    def functor_eval(self, params, rcut, cusp, r, dtype=np.float64):
        params = _np(params, np.float64)
        r = _np(r, dtype)
        u, du, d2u = (np.zeros_like(r) for _ in range(3))
        f = getattr(self.lib, "orc_functor_eval_" + self.suf(dtype))
        self._chk(f(_p(params), C.c_int(len(params)), C.c_double(rcut), C.c_double(cusp), C.c_int(len(r)), _p(r),
                    _p(u), _p(du), _p(d2u)))
        return u, du, d2u

    def functor_coefs(self, params, rcut, cusp, dtype=np.float64):
        params = _np(params, np.float64)
        coefs = np.zeros(len(params) + 4, dtype)
        dri = np.zeros(1, dtype)
        f = getattr(self.lib, "orc_functor_coefs_" + self.suf(dtype))
        self._chk(f(_p(params), C.c_int(len(params)), C.c_double(rcut), C.c_double(cusp), _p(coefs), _p(dri)))
        return coefs, dri[0]

    def dist_row(self, lattice, pos, rsoa, nsrc, flip_ind, dtype=np.float64):
        lattice = _np(lattice, np.float64)
        pos = _np(pos, dtype)
        rsoa = _np(rsoa, dtype)
        npad = rsoa.shape[1]
        out = np.zeros((4, npad), dtype)
        f = getattr(self.lib, "orc_dist_row_" + self.suf(dtype))
        self._chk(f(_p(lattice), _p(pos), _p(rsoa), C.c_long(npad), C.c_int(nsrc), C.c_int(flip_ind), _p(out)))
        return out

    # -- RNG
    def rng_uniform(self, seed, n):
        out = np.zeros(n, np.float64)
        self.lib.orc_rng_uniform(C.c_uint32(seed), C.c_int(n), _p(out))
        return out

    def rng_raw(self, seed, n):
        out = np.zeros(n, np.uint32)
        self.lib.orc_rng_raw(C.c_uint32(seed), C.c_long(n), _p(out))
        return out

    def rng_gauss(self, seed, n, dtype=np.float64):
        out = np.zeros(n, dtype)
        getattr(self.lib, "orc_rng_gauss_" + self.suf(dtype))(C.c_uint32(seed), C.c_int(n), _p(out))
        return out

    def rng(self, seed):
        return OracleRng(self, seed)

    def vmc(self, system, **kw):
        return OracleVMC(self, system, **kw)


class OracleRng:
    """Stateful StdRandom<double> (std::mt19937 + boost-style uniform) with the reference's Box-Muller."""

    def __init__(self, orc, seed):
        self.o = orc
        self.h = C.c_void_p(orc.lib.orc_rng_create(C.c_uint32(seed)))

    def __del__(self):
        try:
            self.o.lib.orc_rng_destroy(self.h)
        except Exception:
            pass

    def uniform(self):
        return self.o.lib.orc_rng_next(self.h)

    def gauss(self, n, dtype=np.float64):
        out = np.zeros(n, dtype)
        getattr(self.o.lib, "orc_rng_gauss_next_" + self.o.suf(dtype))(self.h, C.c_int(n), _p(out))
        return out


class DelayedUpdateHandle:
    def __init__(self, orc, n, k, dtype):
        self.o, self.n, self.k, self.dt = orc, n, k, np.dtype(dtype)
        self.s = orc.suf(dtype)
        self.h = C.c_void_p(getattr(orc.lib, "orc_du_create_" + self.s)(C.c_int(n), C.c_int(k)))

    def __del__(self):
        try:
            getattr(self.o.lib, "orc_du_destroy_" + self.s)(self.h)
        except Exception:
            pass

    def get_inv_row(self, Ainv, row):
        out = np.zeros(self.n, self.dt)
        getattr(self.o.lib, "orc_du_get_inv_row_" + self.s)(self.h, _p(Ainv), C.c_int(Ainv.shape[1]), C.c_int(row), _p(out))
        return out

    def accept_row(self, Ainv, row, psiV, ratio):
        psiV = _np(psiV, self.dt)
        ratio = complex(ratio)
        getattr(self.o.lib, "orc_du_accept_row_" + self.s)(self.h, _p(Ainv), C.c_int(Ainv.shape[1]), C.c_int(row),
                                                          _p(psiV), C.c_double(ratio.real), C.c_double(ratio.imag))

    def pseudo_accept_row(self, Ainv, row):
        getattr(self.o.lib, "orc_du_pseudo_accept_row_" + self.s)(self.h, _p(Ainv), C.c_int(Ainv.shape[1]), C.c_int(row))

    def update_inv_mat(self, Ainv):
        getattr(self.o.lib, "orc_du_update_inv_mat_" + self.s)(self.h, _p(Ainv), C.c_int(Ainv.shape[1]))

    @property
    def delay_count(self):
        return int(getattr(self.o.lib, "orc_du_delay_count_" + self.s)(self.h))


class OracleVMC:
    """The oracle's restatement of VMCBatched::advanceWalkers over a synthetic system (see qmcpack_b200.workload)."""

    def __init__(self, orc, system, nw, ncrowds=1, seeds=None, tau=0.3, use_drift=True, delay_rank=32,
                 batched_engine=True, precision=None, dmc=False):
        self.o = orc
        s = system
        prec = precision if precision is not None else (1 if s["coefs"][0].dtype == np.float32 else 0)
        p = VMCParams()
        p.precision = prec
        p.n_up, p.n_dn = s["n_up"], s["n_dn"]
        p.lattice[:] = list(np.asarray(s["lattice"], np.float64).ravel())
        if s.get("spline_lattice") is not None:  # the primitive cell of the table inside a tiled simulation cell
            p.has_spline_lattice = 1
            p.spline_lattice[:] = list(np.asarray(s["spline_lattice"], np.float64).ravel())
        self._keep = []
        for i in range(2):
            c = s["coefs"][i]
            assert c.flags["C_CONTIGUOUS"]
            p.coefs[i] = c.ctypes.data
            self._keep.append(c)
        p.grid[:] = [int(x) - 3 for x in s["coefs"][0].shape[:3]]
        p.npad = int(s["coefs"][0].shape[3])
        j2 = s.get("j2")
        if j2:
            uu = _np(j2["uu"], np.float64)
            ud = None if j2.get("ud") is None else _np(j2["ud"], np.float64)  # None: the uu functor serves every pair
            self._keep += [uu, ud]
            p.n_j2, p.j2_uu, p.j2_rcut = len(uu), uu.ctypes.data_as(c_dp), j2["rcut"]
            if ud is not None:
                p.j2_ud = ud.ctypes.data_as(c_dp)
        j1 = s.get("j1")
        if j1:
            ip = _np(j1["ion_pos"], np.float64)
            ig = _np(j1["ion_grp"], np.int32)
            prm = _np(j1["params"], np.float64)
            rc = _np(j1["rcut"], np.float64)
            self._keep += [ip, ig, prm, rc]
            p.nions, p.ion_pos, p.ion_grp = len(ig), ip.ctypes.data_as(c_dp), ig.ctypes.data_as(c_ip)
            p.n_ion_groups, p.n_j1 = prm.shape[0], prm.shape[1]
            p.j1_params, p.j1_rcut = prm.ctypes.data_as(c_dp), rc.ctypes.data_as(c_dp)
        p.nw, p.ncrowds = nw, ncrowds
        sd = _np(seeds if seeds is not None else [1000 + c for c in range(ncrowds)], np.uint32)
        self._keep.append(sd)
        p.seeds = sd.ctypes.data_as(C.POINTER(C.c_uint32))
        p.tau, p.use_drift, p.delay_rank = tau, int(use_drift), delay_rank
        p.batched_engine = int(batched_engine and not orc.is_reference)
        # complex orbitals (SplineC2C): system["kpts"] = per-spin [n][3] Cartesian twist vectors, tables hold 2n components
        kp = s.get("kpts")
        self.cplx = kp is not None
        p.complex_orbitals = int(self.cplx)
        p.dmc = int(dmc)
        if self.cplx:
            for i in range(2):
                k = _np(kp[i], np.float64)
                self._keep.append(k)
                p.kpts[i] = k.ctypes.data_as(c_dp)
        self.p = p
        self.N = p.n_up + p.n_dn
        self.nw = nw
        self.h = C.c_void_p(orc.lib.orc_vmc_create(C.byref(p)))
        if not self.h:
            raise RuntimeError(orc.lib.orc_last_error().decode())

    def __del__(self):
        try:
            self.o.lib.orc_vmc_destroy(self.h)
        except Exception:
            pass

    def set_positions(self, R):
        R = _np(R, np.float64)
        assert R.shape == (self.nw, self.N, 3)
        self.o._chk(self.o.lib.orc_vmc_set_positions(self.h, _p(R)))

    def recompute(self):
        self.o._chk(self.o.lib.orc_vmc_recompute(self.h))

    def probe_move(self, iw, iat, displ):
        """evalGrad, makeMove by `displ`, calcRatioGrad of walker iw (nothing accepted): (ratio, grad_old[3], grad_new[3])
        as complex numbers (zero imaginary parts for real orbitals)"""
        d = _np(displ, np.float64)
        r, go, gn = np.zeros(2), np.zeros(6), np.zeros(6)
        self.o._chk(self.o.lib.orc_vmc_probe_move(self.h, C.c_int(iw), C.c_int(iat), _p(d), _p(r), _p(go), _p(gn)))
        return complex(r[0], r[1]), go[0::2] + 1j * go[1::2], gn[0::2] + 1j * gn[1::2]

    def evaluate_ratios(self, iw, ref, r_vp, ct=0):
        """TrialWaveFunction::mw_evaluateRatios of walker iw: psi(electron ref at r_vp[k]) / psi, complex [nk];
        ct 0 ALL, 1 FERMIONIC, 2 NONFERMIONIC"""
        r = _np(r_vp, np.float64).reshape(-1, 3)
        out = np.zeros(2 * len(r))
        self.o._chk(self.o.lib.orc_vmc_evaluate_ratios(self.h, C.c_int(iw), C.c_int(ref), C.c_int(len(r)), _p(r), C.c_int(ct),
                                                       _p(out)))
        return out[0::2] + 1j * out[1::2]

    def sweep(self, nsteps=1, log_accept=False):
        log = np.zeros((nsteps, self.N, self.nw), np.uint8) if log_accept else None
        sec = C.c_double(0)
        self.o._chk(self.o.lib.orc_vmc_sweep(self.h, C.c_int(nsteps), _p(log) if log_accept else None, C.byref(sec)))
        self.last_seconds = sec.value
        return log

    def sweep_forced(self, forced, want_ratios=True):
        forced = _np(forced, np.uint8)
        nsteps = forced.shape[0]
        assert forced.shape == (nsteps, self.N, self.nw)
        ratios = np.zeros((nsteps, self.N, self.nw), np.complex128 if self.cplx else np.float64) if want_ratios else None
        self.o._chk(self.o.lib.orc_vmc_sweep_forced(self.h, C.c_int(nsteps), _p(forced),
                                                    _p(ratios) if want_ratios else None))
        return ratios

    def positions(self):
        R = np.zeros((self.nw, self.N, 3), np.float64)
        self.o._chk(self.o.lib.orc_vmc_get_positions(self.h, _p(R)))
        return R

    def evaluate_gl(self):
        logpsi, ke = np.zeros(self.nw), np.zeros(self.nw)
        dt = np.complex128 if self.cplx else np.float64
        G, L = np.zeros((self.nw, self.N, 3), dt), np.zeros((self.nw, self.N), dt)
        self.o._chk(self.o.lib.orc_vmc_evaluate_gl(self.h, _p(logpsi), _p(ke), _p(G), _p(L)))
        return logpsi, ke, G, L

    def set_num_walkers(self, n):
        self.o._chk(self.o.lib.orc_vmc_set_num_walkers(self.h, C.c_int(n)))
        self.nw = int(n)

    def copy_walker(self, src, dst):
        self.o._chk(self.o.lib.orc_vmc_copy_walker(self.h, C.c_int(src), C.c_int(dst)))

    def rr(self):
        a, p = np.zeros(self.nw), np.zeros(self.nw)
        self.o._chk(self.o.lib.orc_vmc_get_rr(self.h, _p(a), _p(p)))
        return a, p

    def counts(self):
        a, r = np.zeros(self.nw, np.int64), np.zeros(self.nw, np.int64)
        self.o._chk(self.o.lib.orc_vmc_get_counts(self.h, _p(a), _p(r)))
        return a, r

    def psiminv(self, iw, s):
        n = self.p.n_up if s == 0 else self.p.n_dn
        out = np.zeros((n, n), np.complex128 if self.cplx else np.float64)
        ld = C.c_double(0)
        self.o._chk(self.o.lib.orc_vmc_get_psiminv(self.h, C.c_int(iw), C.c_int(s), _p(out), C.byref(ld)))
        return out, ld.value

    def j2_state(self, iw):
        Uat, dUat, d2Uat = np.zeros(self.N), np.zeros((3, self.N)), np.zeros(self.N)
        self.o._chk(self.o.lib.orc_vmc_get_j2(self.h, C.c_int(iw), _p(Uat), _p(dUat), _p(d2Uat)))
        return Uat, dUat, d2Uat


_cache = {}


def port():
    if "port" not in _cache:
        build()
        _cache["port"] = Oracle(PORT_SO)
    return _cache["port"]


def ref():
    """The reference-compiled checker, or None when oracle/_ref has not been built."""
    if "ref" not in _cache:
        build()
        _cache["ref"] = Oracle(REF_SO) if os.path.exists(REF_SO) else None
    return _cache["ref"]
