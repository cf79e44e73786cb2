"""ctypes binding of libqmcb.so (include/qmcb.h) with numpy-facing classes whose method names mirror the reference
interfaces they stand in for:

  SplineSPOSet        SPOSet / BsplineSet / SplineR2R / SplineC2C      mw_evaluateValue, mw_evaluateVGL,
                                                                       mw_evaluateVGLandDetRatioGrads, mw_evaluateDetRatios
  Crowd               Crowd + TrialWaveFunction + ParticleSet batch    mw_recompute, mw_evalGrad, mw_makeMove,
                                                                       mw_calcRatioGrad, mw_accept_rejectMove, ...
  Crowd.det_*         DiracDeterminantBatched / DelayedUpdateBatched   mw_evalGrad, mw_getInvRow, mw_ratioGrad,
                                                                       mw_accept_rejectRow, mw_updateInvMat
  Crowd.j2_*          TwoBodyJastrow                                   mw_ratioGrad, mw_accept_rejectMove

This is the thin test/bench-facing layer; the product is the C ABI.  The library is REQUIRED: importing works without
it (so CPU-only tooling can inspect the package) but every call raises if libqmcb.so or a CUDA device is missing --
there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QMCB_LIB") or os.path.join(HERE, "libqmcb.so")  # QMCB_LIB: tuning variants (scripts/)
FULL, MIXED = 0, 1
R2R, C2C = 0, 1

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_u8p = C.POINTER(C.c_uint8)
vp = C.c_void_p


class QmcbSystem(C.Structure):
    _fields_ = [
        ("precision", C.c_int), ("n_up", C.c_int), ("n_dn", C.c_int),
        ("lattice", C.c_double * 9),
        ("spo", vp * 2), ("delay_rank", C.c_int),
        ("n_j2", C.c_int), ("j2_uu", c_dp), ("j2_ud", c_dp), ("j2_rcut", C.c_double),
        ("nions", C.c_int), ("ion_pos", c_dp), ("ion_grp", c_ip), ("n_ion_groups", C.c_int),
        ("n_j1", C.c_int), ("j1_params", c_dp), ("j1_rcut", c_dp),
    ]


class QmcbVmcParams(C.Structure):
    _fields_ = [("tau", C.c_double), ("use_drift", C.c_int), ("seed", C.c_uint32), ("use_cuda_graph", C.c_int),
                ("dmc", C.c_int), ("sweep_kernel", C.c_int)]


_lib = None

# every symbol include/qmcb.h declares (checked by tests/test_abi.py against the header text)
SYMBOLS = [
    "qmcb_init", "qmcb_last_error", "qmcb_device_count", "qmcb_aligned_size", "qmcb_kernel_launch_count",
    "qmcb_spline_create", "qmcb_spline_destroy", "qmcb_spline_table_bytes", "qmcb_spline_mw_evaluate_value",
    "qmcb_spline_mw_evaluate_vgl", "qmcb_spline_mw_evaluate_vgl_ratio_grads", "qmcb_spline_mw_evaluate_det_ratios",
    "qmcb_spline_mw_vgl_ratio_grads_dev", "qmcb_spline_rg_parts", "qmcb_spline_mw_evaluate_vgl_ratio_grads_offload",
    "qmcb_crowd_create", "qmcb_crowd_destroy", "qmcb_crowd_sync", "qmcb_crowd_device_bytes", "qmcb_crowd_is_complex",
    "qmcb_crowd_set_positions", "qmcb_crowd_get_positions",
    "qmcb_twf_mw_recompute", "qmcb_twf_mw_eval_grad", "qmcb_ps_mw_make_move", "qmcb_twf_mw_calc_ratio_grad",
    "qmcb_twf_mw_accept_reject", "qmcb_twf_mw_complete_updates", "qmcb_twf_mw_evaluate_gl",
    "qmcb_det_mw_eval_grad", "qmcb_det_mw_get_inv_row", "qmcb_det_mw_ratio_grad", "qmcb_det_mw_accept_reject",
    "qmcb_det_mw_complete_updates", "qmcb_det_mw_recompute_from_matrices", "qmcb_det_set_phi_vgl",
    "qmcb_det_mw_ratio_grad_from_phi", "qmcb_det_delay_count", "qmcb_det_time_update_inv_mat", "qmcb_det_time_inverse",
    "qmcb_dtaa_get_temp_rows", "qmcb_j2_mw_ratio_grad", "qmcb_j2_mw_accept_reject", "qmcb_j2_get_state",
    "qmcb_vmc_init", "qmcb_vmc_sweep", "qmcb_vmc_sweep_async", "qmcb_vmc_counts", "qmcb_vmc_sweep_kernel", "qmcb_vmc_profile_sweep", "qmcb_crowd_host_kernel", "qmcb_twf_mw_calc_ratio",
    "qmcb_twf_mw_evaluate_ratios", "qmcb_crowd_stream",
    "qmcb_dmc_get_rr", "qmcb_crowd_walker_bytes", "qmcb_crowd_pack_walker", "qmcb_crowd_unpack_walker",
    "qmcb_crowd_copy_walker", "qmcb_crowd_set_num_walkers", "qmcb_crowd_num_walkers", "qmcb_crowd_capacity",
]


DRIVER_SYMBOLS = ["qmcb_host_vmc_last_error", "qmcb_host_vmc_create", "qmcb_host_vmc_destroy", "qmcb_host_vmc_run",
                  "qmcb_host_vmc_counts", "qmcb_host_vmc_bytes_per_sweep",
                  "qmcb_dmc_last_error", "qmcb_dmc_create", "qmcb_dmc_create_with_engine", "qmcb_dmc_destroy", "qmcb_dmc_step",
                  "qmcb_dmc_advance", "qmcb_dmc_branch", "qmcb_dmc_get_walkers", "qmcb_dmc_set_weights",
                  "qmcb_dmc_branch_weight"]


def lib():
    """Loads libqmcb.so; raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m qmcpack_b200.build` "
                               "(nvcc, sm_100a).  There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.qmcb_last_error.restype = C.c_char_p
        L.qmcb_aligned_size.restype = C.c_size_t
        L.qmcb_aligned_size.argtypes = [C.c_int, C.c_size_t]
        L.qmcb_kernel_launch_count.restype = C.c_ulonglong
        L.qmcb_spline_table_bytes.restype = C.c_size_t
        L.qmcb_spline_table_bytes.argtypes = [vp]
        L.qmcb_crowd_walker_bytes.restype = C.c_size_t
        L.qmcb_crowd_walker_bytes.argtypes = [vp]
        L.qmcb_crowd_device_bytes.restype = C.c_size_t
        L.qmcb_crowd_device_bytes.argtypes = [vp]
        L.qmcb_crowd_stream.restype = vp
        L.qmcb_crowd_stream.argtypes = [vp]
        L.qmcb_spline_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, c_ip, C.c_int, C.c_int, C.c_size_t, vp, c_dp,
                                         c_ip, c_dp]
        L.qmcb_spline_mw_evaluate_vgl_ratio_grads.argtypes = [vp, C.c_int, vp, vp, C.c_size_t, vp, vp, vp]
        L.qmcb_spline_mw_evaluate_det_ratios.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, C.c_size_t, vp]
        L.qmcb_spline_mw_vgl_ratio_grads_dev.argtypes = [vp, C.c_int, vp, vp, C.c_size_t, vp, vp, vp]
        L.qmcb_spline_mw_evaluate_vgl_ratio_grads_offload.argtypes = [vp, C.c_int, vp, vp, C.c_size_t, vp, vp, vp, vp]
        L.qmcb_crowd_create.argtypes = [C.POINTER(vp), C.POINTER(QmcbSystem), C.c_int]
        L.qmcb_det_mw_get_inv_row.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t), vp]
        L.qmcb_host_vmc_last_error.restype = C.c_char_p
        L.qmcb_dmc_last_error.restype = C.c_char_p
        L.qmcb_dmc_branch_weight.restype = C.c_double
        L.qmcb_dmc_branch_weight.argtypes = [vp, C.c_double, C.c_double]
        _lib = L
    return _lib


def _chk(rc):
    if rc != 0:
        raise RuntimeError(lib().qmcb_last_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(vp)


def device_count():
    return int(lib().qmcb_device_count())


def init(device=0):
    _chk(lib().qmcb_init(C.c_int(device)))


def kernel_launch_count():
    return int(lib().qmcb_kernel_launch_count())


def _vt(precision, kind=R2R):
    if kind == C2C:
        return np.complex64 if precision == MIXED else np.complex128
    return np.float32 if precision == MIXED else np.float64


class SplineSPOSet:
    """Tricubic B-spline orbital set resident in HBM (reference: BsplineSet / SplineR2R / SplineC2C)."""

    def __init__(self, coefs, n_orb, G, kind=R2R, halfG=None, kcart=None):
        coefs = np.ascontiguousarray(coefs)
        assert coefs.dtype in (np.float32, np.float64) and coefs.ndim == 4
        self.precision = MIXED if coefs.dtype == np.float32 else FULL
        self.kind = kind
        self.n_orb = int(n_orb)
        self.n_spl = (2 if kind == C2C else 1) * self.n_orb
        self.vt = _vt(self.precision, kind)
        self.st = coefs.dtype
        grid = np.array([s - 3 for s in coefs.shape[:3]], np.int32)
        Gd = np.ascontiguousarray(G, np.float64).reshape(9)
        hg = None if halfG is None else np.ascontiguousarray(halfG, np.int32)
        kc = None if kcart is None else np.ascontiguousarray(kcart, np.float64)
        self.h = vp()
        _chk(lib().qmcb_spline_create(C.byref(self.h), self.precision, kind, grid.ctypes.data_as(c_ip), self.n_orb,
                                      self.n_spl, coefs.shape[3], _p(coefs), Gd.ctypes.data_as(c_dp),
                                      None if hg is None else hg.ctypes.data_as(c_ip),
                                      None if kc is None else kc.ctypes.data_as(c_dp)))

    def __del__(self):
        try:
            if self.h:
                lib().qmcb_spline_destroy(self.h)
        except Exception:
            pass

    @property
    def rg_parts(self):
        return int(lib().qmcb_spline_rg_parts(self.h))

    @property
    def table_bytes(self):
        return int(lib().qmcb_spline_table_bytes(self.h))

    def mw_evaluateValue(self, r):
        r = np.ascontiguousarray(r, np.float64)
        psi = np.zeros((r.shape[0], self.n_orb), self.vt)
        _chk(lib().qmcb_spline_mw_evaluate_value(self.h, C.c_int(r.shape[0]), _p(r), _p(psi)))
        return psi

    def mw_evaluateVGL(self, r):
        r = np.ascontiguousarray(r, np.float64)
        nw = r.shape[0]
        psi = np.zeros((nw, self.n_orb), self.vt)
        dpsi = np.zeros((nw, self.n_orb, 3), self.vt)
        d2psi = np.zeros((nw, self.n_orb), self.vt)
        _chk(lib().qmcb_spline_mw_evaluate_vgl(self.h, C.c_int(nw), _p(r), _p(psi), _p(dpsi), _p(d2psi)))
        return psi, dpsi, d2psi

    def mw_evaluateVGLandDetRatioGrads_offload(self, r, invrow_dev, ld, phi_vgl_dev=None, stream=None):
        """SPOSet::mw_evaluateVGLandDetRatioGrads as an offload SPOSet is driven (DiracDeterminantBatched.cpp:334-346):
        host positions, DEVICE inverse rows [nw][ld] (an address), optional device phi_vgl_v [5][nw][n_orb] (an address);
        returns host ratios [nw] and grads [nw][3]."""
        r = np.ascontiguousarray(r, np.float64)
        nw = r.shape[0]
        ratios = np.zeros(nw, self.vt)
        grads = np.zeros((nw, 3), self.vt)
        _chk(lib().qmcb_spline_mw_evaluate_vgl_ratio_grads_offload(self.h, nw, _p(r), vp(invrow_dev), ld,
                                                                   vp(phi_vgl_dev) if phi_vgl_dev else None, _p(ratios),
                                                                   _p(grads), vp(stream) if stream else None))
        return ratios, grads

    def mw_evaluateVGLandDetRatioGrads(self, r, invrow):
        r = np.ascontiguousarray(r, np.float64)
        nw = r.shape[0]
        invrow = np.ascontiguousarray(invrow, self.vt)
        phi = np.zeros((5, nw, self.n_orb), self.vt)
        ratios = np.zeros(nw, self.vt)
        grads = np.zeros((nw, 3), self.vt)
        _chk(lib().qmcb_spline_mw_evaluate_vgl_ratio_grads(self.h, nw, _p(r), _p(invrow), invrow.shape[1], _p(phi),
                                                           _p(ratios), _p(grads)))
        return phi, ratios, grads

    def mw_evaluateDetRatios(self, r_vp, ref_walker, invrow):
        r_vp = np.ascontiguousarray(r_vp, np.float64)
        ref = np.ascontiguousarray(ref_walker, np.int32)
        invrow = np.ascontiguousarray(invrow, self.vt)
        ratios = np.zeros(r_vp.shape[0], self.vt)
        _chk(lib().qmcb_spline_mw_evaluate_det_ratios(self.h, r_vp.shape[0], _p(r_vp), _p(ref), invrow.shape[0],
                                                      _p(invrow), invrow.shape[1], _p(ratios)))
        return ratios


class Crowd:
    """nw walkers with their trial-wavefunction state on the device; one CUDA stream (reference: Crowd + the
    multi-walker resources of TrialWaveFunction / DiracDeterminantBatched / TwoBodyJastrow / ParticleSet)."""

    def __init__(self, system, nw, delay_rank=32, spo=None):
        s = system
        c0 = s["coefs"][0]
        self.precision = MIXED if c0.dtype == np.float32 else FULL
        self.T = np.float32 if self.precision == MIXED else np.float64
        self.n_up, self.n_dn = int(s["n_up"]), int(s["n_dn"])
        self.N = self.n_up + self.n_dn
        self._nw0 = int(nw)
        self.k = int(delay_rank)
        lat = np.ascontiguousarray(s["lattice"], np.float64).reshape(3, 3)
        G = np.linalg.inv(lat)  # CrystalLattice: G = inverse(R), ru = r . G
        kp = s.get("kpts")
        if spo is None:
            if kp is not None:  # complex orbitals: SplineC2C tables + twist vectors
                up = SplineSPOSet(s["coefs"][0], self.n_up, G, kind=C2C, kcart=kp[0])
                dn = SplineSPOSet(s["coefs"][1], self.n_dn, G, kind=C2C, kcart=kp[1])
            else:
                up = SplineSPOSet(s["coefs"][0], self.n_up, G)
                dn = up if s["coefs"][1] is s["coefs"][0] and self.n_dn == self.n_up else SplineSPOSet(s["coefs"][1], self.n_dn, G)
            spo = (up, dn)
        self.spo = spo
        self.cplx = spo[0].kind == C2C
        # V: determinant value type (inverse rows, orbital rows, ratios of the component-level calls);
        # P: PsiValue / gradient type of the trial-wavefunction-level calls (always double precision)
        self.V = _vt(self.precision, C2C if self.cplx else R2R)
        self.P = np.complex128 if self.cplx else np.float64
        q = QmcbSystem()
        q.precision, q.n_up, q.n_dn = self.precision, self.n_up, self.n_dn
        q.lattice[:] = list(lat.ravel())
        q.spo[0], q.spo[1] = spo[0].h.value, spo[1].h.value
        q.delay_rank = self.k
        self._keep = []
        j2 = s.get("j2")
        if j2:
            uu = np.ascontiguousarray(j2["uu"], np.float64)
            # ud = None: only the like-spin correlation was given and serves every pair (TwoBodyJastrow::addFunc)
            ud = None if j2.get("ud") is None else np.ascontiguousarray(j2["ud"], np.float64)
            self._keep += [uu, ud]
            q.n_j2, q.j2_uu, q.j2_rcut = len(uu), uu.ctypes.data_as(c_dp), j2["rcut"]
            if ud is not None:
                q.j2_ud = ud.ctypes.data_as(c_dp)
        j1 = s.get("j1")
        if j1:
            ip = np.ascontiguousarray(j1["ion_pos"], np.float64)
            ig = np.ascontiguousarray(j1["ion_grp"], np.int32)
            prm = np.ascontiguousarray(j1["params"], np.float64)
            rc = np.ascontiguousarray(j1["rcut"], np.float64)
            self._keep += [ip, ig, prm, rc]
            q.nions, q.ion_pos, q.ion_grp = len(ig), ip.ctypes.data_as(c_dp), ig.ctypes.data_as(c_ip)
            q.n_ion_groups, q.n_j1 = prm.shape[0], prm.shape[1]
            q.j1_params, q.j1_rcut = prm.ctypes.data_as(c_dp), rc.ctypes.data_as(c_dp)
        self.h = vp()
        _chk(lib().qmcb_crowd_create(C.byref(self.h), C.byref(q), self._nw0))

    def __del__(self):
        try:
            if self.h:
                lib().qmcb_crowd_destroy(self.h)
        except Exception:
            pass

    def n_of(self, spin):
        return self.n_up if spin == 0 else self.n_dn

    @property
    def device_bytes(self):
        return int(lib().qmcb_crowd_device_bytes(self.h))

    @property
    def stream(self):
        return lib().qmcb_crowd_stream(self.h)

    def sync(self):
        _chk(lib().qmcb_crowd_sync(self.h))

    # ---- ParticleSet
    def set_positions(self, R):
        R = np.ascontiguousarray(R, np.float64)
        assert R.shape == (self.nw, self.N, 3)
        _chk(lib().qmcb_crowd_set_positions(self.h, _p(R)))

    def positions(self):
        R = np.zeros((self.nw, self.N, 3))
        _chk(lib().qmcb_crowd_get_positions(self.h, _p(R)))
        return R

    def mw_makeMove(self, iat, displ):
        d = np.ascontiguousarray(displ, np.float64)
        assert d.shape == (self.nw, 3)
        _chk(lib().qmcb_ps_mw_make_move(self.h, C.c_int(iat), _p(d)))

    # ---- TrialWaveFunction
    def mw_recompute(self):
        _chk(lib().qmcb_twf_mw_recompute(self.h))

    def mw_evalGrad(self, iat):
        g = np.zeros((self.nw, 3), self.P)
        _chk(lib().qmcb_twf_mw_eval_grad(self.h, C.c_int(iat), _p(g)))
        return g

    def mw_calcRatioGrad(self, iat):
        r, g = np.zeros(self.nw, self.P), np.zeros((self.nw, 3), self.P)
        _chk(lib().qmcb_twf_mw_calc_ratio_grad(self.h, C.c_int(iat), _p(r), _p(g)))
        return r, g

    def mw_accept_rejectMove(self, iat, accepted, safe_to_delay=True):
        a = np.ascontiguousarray(accepted, np.uint8)
        assert a.shape == (self.nw,)
        _chk(lib().qmcb_twf_mw_accept_reject(self.h, C.c_int(iat), _p(a), C.c_int(1 if safe_to_delay else 0)))

    def mw_completeUpdates(self):
        _chk(lib().qmcb_twf_mw_complete_updates(self.h))

    def mw_evaluateGL(self):
        G, L = np.zeros((self.nw, self.N, 3), self.P), np.zeros((self.nw, self.N), self.P)
        lp, ke = np.zeros(self.nw), np.zeros(self.nw)
        _chk(lib().qmcb_twf_mw_evaluate_gl(self.h, _p(G), _p(L), _p(lp), _p(ke)))
        return lp, ke, G, L

    def mw_calcRatio(self, iat):
        """TrialWaveFunction::mw_calcRatio: ratio of the proposed move without gradients"""
        r = np.zeros(self.nw, self.P)
        _chk(lib().qmcb_twf_mw_calc_ratio(self.h, C.c_int(iat), _p(r)))
        return r

    def mw_evaluateRatios(self, walker, ref_ptcl, r_vp, compute_type=0):
        """TrialWaveFunction::mw_evaluateRatios at virtual positions (NLPP quadrature points): walker [nvp], ref_ptcl [nvp],
        r_vp [nvp][3]; compute_type 0 ALL, 1 FERMIONIC, 2 NONFERMIONIC"""
        wk = np.ascontiguousarray(walker, np.int32)
        rf = np.ascontiguousarray(ref_ptcl, np.int32)
        r = np.ascontiguousarray(r_vp, np.float64).reshape(-1, 3)
        out = np.zeros(len(wk), self.P)
        _chk(lib().qmcb_twf_mw_evaluate_ratios(self.h, C.c_int(len(wk)), _p(wk), _p(rf), _p(r), C.c_int(compute_type), _p(out)))
        return out

    def mw_block_estimators(self):
        """(log psi, kinetic energy) per walker without shipping G and L to the host: what a block estimator needs"""
        lp, ke = np.zeros(self.nw), np.zeros(self.nw)
        _chk(lib().qmcb_twf_mw_evaluate_gl(self.h, None, None, _p(lp), _p(ke)))
        return lp, ke

    # ---- DiracDeterminantBatched / DelayedUpdateBatched
    def det_mw_evalGrad(self, spin, row):
        g = np.zeros((self.nw, 3), self.V)
        _chk(lib().qmcb_det_mw_eval_grad(self.h, C.c_int(spin), C.c_int(row), _p(g)))
        return g

    def det_mw_getInvRow(self, spin, row):
        out = np.zeros((self.nw, self.n_of(spin)), self.V)
        dev, ld = vp(), C.c_size_t()
        _chk(lib().qmcb_det_mw_get_inv_row(self.h, spin, row, C.byref(dev), C.byref(ld), _p(out)))
        return out

    def det_mw_ratioGrad(self, spin, row, from_phi=False):
        r, g = np.zeros(self.nw, self.V), np.zeros((self.nw, 3), self.V)
        f = lib().qmcb_det_mw_ratio_grad_from_phi if from_phi else lib().qmcb_det_mw_ratio_grad
        _chk(f(self.h, C.c_int(spin), C.c_int(row), _p(r), _p(g)))
        return r, g

    def det_mw_accept_rejectRow(self, spin, row, accepted):
        a = np.ascontiguousarray(accepted, np.uint8)
        _chk(lib().qmcb_det_mw_accept_reject(self.h, C.c_int(spin), C.c_int(row), _p(a)))

    def det_mw_completeUpdates(self, spin):
        n = self.n_of(spin)
        inv = np.zeros((self.nw, n, n), self.V)
        ld = np.zeros((self.nw, 2))
        _chk(lib().qmcb_det_mw_complete_updates(self.h, C.c_int(spin), _p(inv), _p(ld)))
        return inv, ld

    def det_recompute_from_matrices(self, spin, psiM, dpsiM=None, d2psiM=None):
        n = self.n_of(spin)
        psiM = np.ascontiguousarray(psiM, self.V)
        assert psiM.shape == (self.nw, n, n)
        dp = None if dpsiM is None else np.ascontiguousarray(dpsiM, self.V)
        d2 = None if d2psiM is None else np.ascontiguousarray(d2psiM, self.V)
        _chk(lib().qmcb_det_mw_recompute_from_matrices(self.h, C.c_int(spin), _p(psiM), _p(dp), _p(d2)))

    def det_set_phi_vgl(self, spin, phi):
        phi = np.ascontiguousarray(phi, self.V)
        assert phi.shape == (5, self.nw, self.n_of(spin))
        _chk(lib().qmcb_det_set_phi_vgl(self.h, C.c_int(spin), _p(phi)))

    def det_delay_count(self, spin):
        return int(lib().qmcb_det_delay_count(self.h, C.c_int(spin)))

    def det_time_inverse(self, spin, method, reps=3):
        """microseconds per FP64 inverse + log-determinant of the whole batch (1: cuBLAS getrf/getriBatched, 2: own kernels)"""
        us = C.c_double(0.0)
        _chk(lib().qmcb_det_time_inverse(self.h, C.c_int(spin), C.c_int(method), C.c_int(reps), C.byref(us)))
        return us.value

    def det_time_update_inv_mat(self, spin, delay_count, reps=10):
        """microseconds per mw_updateInvMat launch (measurement hook; recompute the crowd afterwards)"""
        us = C.c_double(0.0)
        _chk(lib().qmcb_det_time_update_inv_mat(self.h, C.c_int(spin), C.c_int(delay_count), C.c_int(reps), C.byref(us)))
        return us.value

    # ---- distance rows / TwoBodyJastrow
    def dtaa_temp_rows(self):
        rows = np.zeros((2, self.nw, 4, self.N), self.T)
        _chk(lib().qmcb_dtaa_get_temp_rows(self.h, _p(rows)))
        return rows

    def j2_mw_ratioGrad(self, iat):
        r, g = np.zeros(self.nw), np.zeros((self.nw, 3), self.T)
        _chk(lib().qmcb_j2_mw_ratio_grad(self.h, C.c_int(iat), _p(r), _p(g)))
        return r, g

    def j2_mw_accept_rejectMove(self, iat, accepted):
        a = np.ascontiguousarray(accepted, np.uint8)
        _chk(lib().qmcb_j2_mw_accept_reject(self.h, C.c_int(iat), _p(a)))

    def j2_state(self, iw):
        U, dU, d2U = np.zeros(self.N), np.zeros((3, self.N)), np.zeros(self.N)
        _chk(lib().qmcb_j2_get_state(self.h, C.c_int(iw), _p(U), _p(dU), _p(d2U)))
        return U, dU, d2U

    # ---- device-resident VMC driver
    def vmc_init(self, tau=0.3, use_drift=True, seed=1000, use_cuda_graph=True, dmc=False, sweep_kernel=0):
        """dmc=True: the move loop of DMCBatched::advanceWalkers (phase rejection, rr accumulators).
        sweep_kernel: 0 automatic, 1 two-kernel path, 2 persistent walker-segment kernel (error if not eligible)"""
        p = QmcbVmcParams(tau, int(use_drift), seed, int(use_cuda_graph), int(dmc), int(sweep_kernel))
        _chk(lib().qmcb_vmc_init(self.h, C.byref(p)))

    @property
    def host_kernel(self):
        """2: per-electron calls served by a resident walker-segment kernel (mailboxes), 1: launches per call, 0: none yet"""
        return int(lib().qmcb_crowd_host_kernel(self.h))

    def vmc_profile_sweep(self):
        """one sweep with an event pair around every launch: dict of sweep time and per-kernel sums (us) and launch counts"""
        out = np.zeros(9)
        _chk(lib().qmcb_vmc_profile_sweep(self.h, _p(out)))
        names = ("segment", "boundary", "gather", "flush")
        d = {"sweep_us": out[0]}
        for i, nm in enumerate(names):
            d[nm + "_us"], d[nm + "_launches"] = out[1 + i], int(out[5 + i])
        return d

    @property
    def sweep_kernel(self):
        """2: persistent walker-segment kernel, 1: boundary + gather kernels per move, 0: driver not initialised"""
        return int(lib().qmcb_vmc_sweep_kernel(self.h))

    # ---- the engine interface of qmcpack_b200.dmc.DMC
    def dmc_sweep(self):
        self.last_log = self.vmc_sweep(1, log_accept=True)

    def local_energies(self):
        return self.mw_evaluateGL()[1]

    def rr(self):
        return self.dmc_rr()

    def dmc_rr(self):
        a, p = np.zeros(self.nw), np.zeros(self.nw)
        _chk(lib().qmcb_dmc_get_rr(self.h, _p(a), _p(p)))
        return a, p

    # ---- walker state for branching / load balancing (device-resident packed buffers: pass raw device pointers,
    #      e.g. torch_tensor.data_ptr() of a uint8 tensor of walker_bytes elements)
    @property
    def walker_bytes(self):
        return int(lib().qmcb_crowd_walker_bytes(self.h))

    def pack_walker(self, iw, dev_ptr):
        _chk(lib().qmcb_crowd_pack_walker(self.h, C.c_int(iw), C.c_void_p(dev_ptr)))

    def unpack_walker(self, iw, dev_ptr):
        _chk(lib().qmcb_crowd_unpack_walker(self.h, C.c_int(iw), C.c_void_p(dev_ptr)))

    def copy_walker(self, src, dst):
        _chk(lib().qmcb_crowd_copy_walker(self.h, C.c_int(src), C.c_int(dst)))

    def set_num_walkers(self, n):
        """live walkers [0, n) of the capacity given at creation (DMC population changes)"""
        _chk(lib().qmcb_crowd_set_num_walkers(self.h, C.c_int(n)))

    @property
    def nw(self):
        """live walkers (the C++ DMC layer changes the count behind this wrapper's back)"""
        h = getattr(self, "h", None)
        return int(lib().qmcb_crowd_num_walkers(h)) if h else self._nw0

    @property
    def capacity(self):
        return int(lib().qmcb_crowd_capacity(self.h))

    def vmc_sweep(self, nsteps=1, log_accept=False):
        log = np.zeros((nsteps, self.N, self.nw), np.uint8) if log_accept else None
        _chk(lib().qmcb_vmc_sweep(self.h, C.c_int(nsteps), _p(log)))
        return log

    def vmc_sweep_async(self):
        _chk(lib().qmcb_vmc_sweep_async(self.h))

    def vmc_counts(self):
        a, r = np.zeros(self.nw, np.int64), np.zeros(self.nw, np.int64)
        _chk(lib().qmcb_vmc_counts(self.h, _p(a), _p(r)))
        return a, r


class HostVMC:
    """The compiled host driver above the C ABI (include/qmcb_driver.h): VMCBatched::advanceWalkers with one host thread
    per crowd, host buffers crossing PCIe every move."""

    def __init__(self, crowds, seeds, tau=0.3, use_drift=True):
        self.crowds = list(crowds)
        nc = len(self.crowds)
        arr = (vp * nc)(*[c.h.value for c in self.crowds])
        nws = np.array([c.nw for c in self.crowds], np.int32)
        sd = np.ascontiguousarray(seeds, np.uint32)
        assert len(sd) == nc
        self.N = self.crowds[0].N
        self.nw_total = int(nws.sum())
        self.h = vp()
        rc = lib().qmcb_host_vmc_create(C.byref(self.h), arr, _p(nws), C.c_int(nc), C.c_int(self.N),
                                        C.c_int(self.crowds[0].precision), _p(sd), C.c_double(tau),
                                        C.c_int(int(use_drift)))
        if rc:
            raise RuntimeError(lib().qmcb_host_vmc_last_error().decode())

    def __del__(self):
        try:
            if self.h:
                lib().qmcb_host_vmc_destroy(self.h)
        except Exception:
            pass

    def run(self, nsteps=1, log_accept=False):
        log = np.zeros((nsteps, self.N, self.nw_total), np.uint8) if log_accept else None
        rc = lib().qmcb_host_vmc_run(self.h, C.c_int(nsteps), _p(log))
        if rc:
            raise RuntimeError(lib().qmcb_host_vmc_last_error().decode())
        return log

    def counts(self):
        a, r = C.c_longlong(), C.c_longlong()
        lib().qmcb_host_vmc_counts(self.h, C.byref(a), C.byref(r))
        return a.value, r.value

    def bytes_per_sweep(self):
        a, r = C.c_longlong(), C.c_longlong()
        lib().qmcb_host_vmc_bytes_per_sweep(self.h, C.byref(a), C.byref(r))
        return a.value, r.value


# ---------------------------------------------------------------------------------------------------------------------
# DMC layer (include/qmcb_driver.h, csrc/dmc_host.cpp): DMCBatched::advanceWalkers + WalkerControl::branch + SFNBranch in
# C++; this class only marshals the communicator (torch.distributed) and, for tests, a duck-typed engine
# ---------------------------------------------------------------------------------------------------------------------
_ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, vp, c_dp, C.c_int)
_SEND_FN = C.CFUNCTYPE(C.c_int, vp, C.c_int, c_dp)
_RECV_FN = C.CFUNCTYPE(C.c_int, vp, C.c_int, c_dp)


class QmcbComm(C.Structure):
    _fields_ = [("rank", C.c_int), ("size", C.c_int), ("ctx", vp), ("allreduce_sum", _ALLREDUCE_FN), ("send", _SEND_FN),
                ("recv", _RECV_FN), ("send_buf", vp), ("recv_buf", vp)]


_E_INT = C.CFUNCTYPE(C.c_int, vp)
_E_DP = C.CFUNCTYPE(C.c_int, vp, c_dp)
_E_DP2 = C.CFUNCTYPE(C.c_int, vp, c_dp, c_dp)
_E_II = C.CFUNCTYPE(C.c_int, vp, C.c_int, C.c_int)
_E_I = C.CFUNCTYPE(C.c_int, vp, C.c_int)
_E_IP = C.CFUNCTYPE(C.c_int, vp, C.c_int, vp)


class QmcbDmcEngine(C.Structure):
    _fields_ = [("ctx", vp), ("num_walkers", _E_INT), ("capacity", _E_INT), ("sweep", _E_INT), ("local_energies", _E_DP),
                ("get_rr", _E_DP2), ("copy_walker", _E_II), ("set_num_walkers", _E_I), ("pack_walker", _E_IP),
                ("unpack_walker", _E_IP)]


class QmcbDmcParams(C.Structure):
    _fields_ = [("tau", C.c_double), ("target_walkers", C.c_int), ("branch_seed", C.c_uint32), ("sigma2", C.c_double),
                ("sigma_bound", C.c_double), ("feedback", C.c_double), ("warmup_steps", C.c_int),
                ("energy_update_interval", C.c_int), ("use_tau_eff", C.c_int)]


class QmcbDmcEnsemble(C.Structure):
    _fields_ = [("energy", C.c_double), ("variance", C.c_double), ("weight", C.c_double), ("num_samples", C.c_double),
                ("r2_accepted", C.c_double), ("r2_proposed", C.c_double), ("living_fraction", C.c_double),
                ("e_trial", C.c_double), ("e_ref", C.c_double), ("tau_eff", C.c_double), ("branch_cutoff", C.c_double),
                ("population", C.c_int), ("local", C.c_int), ("walkers_sent", C.c_longlong),
                ("walkers_received", C.c_longlong), ("bytes_sent", C.c_longlong), ("bytes_received", C.c_longlong)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class TorchComm:
    """qmcb_comm over torch.distributed: NCCL on the GPU box (all-reduce of curData, ncclSend / ncclRecv of the packed
    device-resident walker), gloo in the CPU tests.  Owns the two transfer buffers."""

    def __init__(self, dist, walker_bytes, device):
        import torch
        self.dist, self.device = dist, device
        self.rank, self.size = dist.get_rank(), dist.get_world_size()
        self.send_buf = torch.empty(max(16, walker_bytes), dtype=torch.uint8, device=device)
        self.recv_buf = torch.empty_like(self.send_buf)
        self.bytes_sent = self.bytes_received = 0
        self.messages = 0

        def allreduce(ctx, buf, n):
            try:
                a = np.ctypeslib.as_array(buf, shape=(n,))
                t = torch.as_tensor(a.copy(), dtype=torch.float64, device=device)
                dist.all_reduce(t)
                a[:] = t.cpu().numpy()
                return 0
            except Exception:  # surfaced as "allreduce failed" by the C++ side
                import traceback
                traceback.print_exc()
                return 1

        def send(ctx, dst, header):
            try:
                h = torch.as_tensor(np.ctypeslib.as_array(header, shape=(4,)).copy(), dtype=torch.float64, device=device)
                dist.send(h, dst=dst)
                dist.send(self.send_buf, dst=dst)
                self.bytes_sent += self.send_buf.numel() + 32
                self.messages += 1
                return 0
            except Exception:
                import traceback
                traceback.print_exc()
                return 1

        def recv(ctx, src, header):
            try:
                h = torch.empty(4, dtype=torch.float64, device=device)
                dist.recv(h, src=src)
                dist.recv(self.recv_buf, src=src)
                np.ctypeslib.as_array(header, shape=(4,))[:] = h.cpu().numpy()
                self.bytes_received += self.recv_buf.numel() + 32
                return 0
            except Exception:
                import traceback
                traceback.print_exc()
                return 1

        self._cb = (_ALLREDUCE_FN(allreduce), _SEND_FN(send), _RECV_FN(recv))
        self.c = QmcbComm(self.rank, self.size, None, self._cb[0], self._cb[1], self._cb[2], self.send_buf.data_ptr(),
                          self.recv_buf.data_ptr())


class DMCDriver:
    """The C++ DMC layer on one crowd per rank (or, for tests, on a duck-typed engine with nw, capacity, dmc_sweep(),
    local_energies(), rr(), copy_walker(), set_num_walkers(), pack_walker(), unpack_walker())."""

    def __init__(self, engine, tau, target_walkers=0, branch_seed=7, comm=None, warmup_steps=1 << 30,
                 energy_update_interval=1, sigma2=10.0, sigma_bound=10.0, feedback=1.0, use_tau_eff=True):
        self.engine, self.comm = engine, comm
        self.p = QmcbDmcParams(tau, int(target_walkers), int(branch_seed), sigma2, sigma_bound, feedback, int(warmup_steps),
                               int(energy_update_interval), int(use_tau_eff))
        self.h = vp()
        cp = C.byref(comm.c) if comm is not None else None
        if isinstance(engine, Crowd):
            rc = lib().qmcb_dmc_create(C.byref(self.h), engine.h, C.byref(self.p), cp)
        else:
            self._eng_c = self._wrap_engine(engine)
            rc = lib().qmcb_dmc_create_with_engine(C.byref(self.h), C.byref(self._eng_c), C.byref(self.p), cp)
        if rc:
            raise RuntimeError(lib().qmcb_dmc_last_error().decode())
        self.history = []

    def _wrap_engine(self, e):
        def guard(f):
            def g(*a):
                try:
                    f(*a)
                    return 0
                except Exception:
                    import traceback
                    traceback.print_exc()
                    return 1
            return g

        def energies(ctx, out):
            np.ctypeslib.as_array(out, shape=(e.nw,))[:] = np.asarray(e.local_energies(), np.float64)

        def rr(ctx, a, p):
            ra, rp = e.rr()
            np.ctypeslib.as_array(a, shape=(e.nw,))[:] = np.asarray(ra, np.float64)
            np.ctypeslib.as_array(p, shape=(e.nw,))[:] = np.asarray(rp, np.float64)

        self._eng_cb = (_E_INT(lambda ctx: int(e.nw)), _E_INT(lambda ctx: int(e.capacity)),
                        _E_INT(guard(lambda ctx: e.dmc_sweep())), _E_DP(guard(energies)), _E_DP2(guard(rr)),
                        _E_II(guard(lambda ctx, s, d: e.copy_walker(s, d))),
                        _E_I(guard(lambda ctx, n: e.set_num_walkers(n))),
                        _E_IP(guard(lambda ctx, iw, buf: e.pack_walker(iw, buf))),
                        _E_IP(guard(lambda ctx, iw, buf: e.unpack_walker(iw, buf))))
        return QmcbDmcEngine(None, *self._eng_cb)

    def __del__(self):
        try:
            if self.h:
                lib().qmcb_dmc_destroy(self.h)
        except Exception:
            pass

    def _ens(self, rc, ens):
        if rc:
            raise RuntimeError(lib().qmcb_dmc_last_error().decode())
        d = ens.as_dict()
        self.history.append(d)
        return d

    def step(self, it):
        """one generation: advance + branch(iter, do_not_branch = (iter == 0)) + updateParamAfterPopControl"""
        ens = QmcbDmcEnsemble()
        return self._ens(lib().qmcb_dmc_step(self.h, C.c_int(it), C.byref(ens)), ens)

    def advance(self):
        if lib().qmcb_dmc_advance(self.h):
            raise RuntimeError(lib().qmcb_dmc_last_error().decode())

    def branch_step(self, it=1, do_not_branch=False):
        ens = QmcbDmcEnsemble()
        return self._ens(lib().qmcb_dmc_branch(self.h, C.c_int(it), C.c_int(int(do_not_branch)), C.byref(ens)), ens)

    def walkers(self):
        """(weights, energies, ages) of the live walkers"""
        cap = 1 << 16
        w, e, a = np.zeros(cap), np.zeros(cap), np.zeros(cap, np.int64)
        n = lib().qmcb_dmc_get_walkers(self.h, _p(w), _p(e), _p(a), C.c_int(cap))
        return w[:n].copy(), e[:n].copy(), a[:n].copy()

    def set_weights(self, w):
        w = np.ascontiguousarray(w, np.float64)
        if lib().qmcb_dmc_set_weights(self.h, _p(w), C.c_int(len(w))):
            raise RuntimeError(lib().qmcb_dmc_last_error().decode())

    def branch_weight(self, enew, eold):
        return float(lib().qmcb_dmc_branch_weight(self.h, C.c_double(enew), C.c_double(eold)))
