"""The trial wavefunction (Slater determinants over spline SPOs + two-body Jastrow) on the reference's own diamondC 2x1x1
data, against the literals of QMCWaveFunctions/tests/test_TrialWaveFunction_diamondC_2x1x1.cpp (QMC_COMPLEX branch:
SplineC2C orbitals from two primitive-cell twists, complex determinants, meshfactor 1.5, double tables; 4 electrons in the
2x1x1 supercell; J2 with only the u-u correlation given, which TwoBodyJastrow::addFunc then uses for every pair, rcut =
the supercell's Wigner-Seitz radius).  The test drives two walkers through mw_evaluateLog, mw_evalGrad, mw_makeMove,
mw_calcRatioGrad and mw_accept_rejectMove (:259-405) -- exactly the calls of the C ABI.

The same sequence on the 1x1x1 primitive cell, test_TrialWaveFunction.cpp:56-360, pins the REAL branch (SplineR2R, real
determinants, precision="float" table, Gamma point).

CPU: the wavefunction is restated with numpy on top of the oracle's spline and functor primitives (explicit 2x2
determinants), which pins those primitives AND documents what the literals mean.  GPU: the same sequence through
qmcb_twf_mw_* on a two-walker crowd."""
import itertools
import os

import numpy as np
import pytest

import eshdf_spline
import oracle_lib

HERE = os.path.dirname(os.path.abspath(__file__))
R_SUPER = np.array([[6.7463223, 6.7463223, 0.0], [0.0, 3.37316115, 3.37316115], [3.37316115, 0.0, 3.37316115]])
UU = [0.02904699284, -0.1004179, -0.1752703883, -0.2232576505, -0.2728029201, -0.3253286875, -0.3624525145, -0.3958223107,
      -0.4268582166, -0.4394531176]
R0 = np.array([[0.0, 0.0, 0.0], [0.0, 1.0, 1.0], [1.0, 1.0, 0.0], [1.0, 0.0, 1.0]])  # :88-91
DELTA = np.array([0.1, 0.1, 0.2])           # :212
DELTA_SIGN = np.array([0.1, 0.1, -0.2])     # :310
SHIFTS = np.array([np.array(c) @ R_SUPER for c in itertools.product(range(-2, 3), repeat=3)])
WS_RADIUS = min(np.linalg.norm(s) for s in SHIFTS if np.any(s)) / 2  # default rcut of a periodic J2

# literals (QMC_COMPLEX branch)
LOGPSI_0, LOGPSI_MOVED = -4.546410485374186, -6.626861768296886          # :183, :237
RATIO_ALL, RATIO_FERMI, RATIO_BOSE = 0.1248738460467855, 0.1362181543980086, 0.9167195562048454  # :224-231
GRAD_OLD = [[complex(713.71203320653, 0.020838031926441), complex(713.71203320654, 0.020838031928415),
             complex(-768.42842826889, -0.020838032018344)],
            [complex(118.02653358655, -0.0022419843505538), complex(118.02653358655, -0.0022419843498631),
             complex(-118.46325895634, 0.0022419843493758)]]           # :293-304
R_SIGN = [complex(253.71869245791, -0.00034808849808193), complex(36.915636007059, -6.4240180082292e-05)]  # :325-326
GRAD_SIGN_1 = [complex(1.4567170375539, 0.00027263382943948), complex(1.4567170375539, 0.00027263382945093),
               complex(-1.2930978490431, -0.00027378452214318)]       # :327-329


@pytest.fixture(scope="module")
def orc():
    oracle_lib.build()
    return oracle_lib.port()


@pytest.fixture(scope="module")
def spo_data(orc):
    d = np.load(os.path.join(HERE, "golden", "diamondC_2x1x1_eshdf.npz"))
    Gp = np.linalg.inv(d["primitive_vectors"])
    tw = [int(k) for k, _ in d["band_labels"][:2]]  # size="2": (Gamma, band 0) and ((1/2,0,0), band 0)
    coefs = eshdf_spline.build_table_c2c_twists(orc, d["psi_g"][:2], tw, d["reduced_k"], d["gvectors"], np.float64,
                                                meshfactor=1.5)
    assert coefs.shape[:3] == (69, 63, 63)  # mesh 66 x 60 x 60
    kc = np.stack([eshdf_spline.k_cart(Gp, d["reduced_k"][k]) for k in tw])
    return coefs, Gp, kc


class NumpyTWF:
    """psi = det(up) det(dn) exp(-sum_{i<j} u(r_ij)), restated on the oracle's primitives"""

    def __init__(self, orc, spo_data):
        self.orc = orc
        self.coefs, self.Gp, self.kc = spo_data

    def spo(self, r):
        psi, dpsi, _ = self.orc.c2c_vgl(self.coefs, self.Gp, self.kc, 2, np.atleast_2d(r))
        return psi.astype(np.complex128), dpsi.astype(np.complex128)

    def u(self, r):
        return [x[0] for x in self.orc.functor_eval(UU, WS_RADIUS, -0.25, np.array([r]))]

    @staticmethod
    def min_image(dv):
        c = dv + SHIFTS
        return c[np.argmin((c**2).sum(1))]

    def log_det(self, R):
        psi, _ = self.spo(R)
        return np.log(np.linalg.det(psi[:2]) + 0j) + np.log(np.linalg.det(psi[2:]) + 0j)

    def log_j2(self, R):
        return -sum(self.u(np.linalg.norm(self.min_image(R[i] - R[j])))[0] for i in range(4) for j in range(i))

    def logpsi(self, R):
        return self.log_det(R) + self.log_j2(R)

    def grad(self, R, i):
        grp = slice(0, 2) if i < 2 else slice(2, 4)
        psi, dpsi = self.spo(R[grp])
        ainv = np.linalg.inv(psi)
        e = i - grp.start
        g = np.array([sum(ainv[j][e] * dpsi[e][j][d] for j in range(2)) for d in range(3)])
        for j in range(4):
            if j != i:
                dv = self.min_image(R[j] - R[i])
                r = np.linalg.norm(dv)
                g = g + self.u(r)[1] * dv / r  # grad_i of -u(r_ij) = u'(r) (r_j - r_i) / r
        return g


def test_numpy_restatement_reproduces_the_reference_literals(orc, spo_data):
    w = NumpyTWF(orc, spo_data)
    assert WS_RADIUS == pytest.approx(2.3851851232, rel=1e-9)
    l0 = w.logpsi(R0)
    assert l0.real == pytest.approx(LOGPSI_0, rel=1e-10)
    assert l0.imag == pytest.approx(-3.141586279080522, rel=1e-9)  # :277
    R1 = R0.copy()
    R1[0] += DELTA
    l1 = w.logpsi(R1)
    assert l1.real == pytest.approx(LOGPSI_MOVED, rel=1e-9)
    assert np.exp(l1 - l0) == pytest.approx(RATIO_ALL, rel=1e-9)
    assert np.exp(w.log_det(R1) - w.log_det(R0)) == pytest.approx(RATIO_FERMI, rel=1e-9)
    assert np.exp(w.log_j2(R1) - w.log_j2(R0)) == pytest.approx(RATIO_BOSE, rel=1e-9)
    assert w.grad(R1, 0) == pytest.approx(GRAD_OLD[0], rel=1e-8)
    assert w.grad(R0, 0) == pytest.approx(GRAD_OLD[1], rel=1e-8)
    R1s, R0s = R1.copy(), R0.copy()
    R1s[0] += DELTA_SIGN
    R0s[0] += DELTA_SIGN
    assert np.exp(w.logpsi(R1s) - l1) == pytest.approx(R_SIGN[0], rel=1e-8)
    assert np.exp(w.logpsi(R0s) - l0) == pytest.approx(R_SIGN[1], rel=1e-8)
    assert w.grad(R0s, 0) == pytest.approx(GRAD_SIGN_1, rel=1e-7)


@pytest.mark.gpu
def test_gpu_twf_sequence_matches_the_reference_literals(orc, spo_data):
    from qmcpack_b200 import api, build
    build.build()
    api.init(0)
    coefs, Gp, kc = spo_data
    up = api.SplineSPOSet(coefs, 2, Gp, kind=api.C2C, kcart=kc)  # the spline lives on the PRIMITIVE cell
    system = dict(n_up=2, n_dn=2, lattice=R_SUPER, coefs=[coefs, coefs], kpts=[kc, kc],
                  j2=dict(uu=UU, ud=None, rcut=WS_RADIUS))
    crowd = api.Crowd(system, nw=2, delay_rank=2, spo=(up, up))
    R1 = R0.copy()
    R1[0] += DELTA
    crowd.set_positions(np.stack([R1, R0]))  # walker 0 already moved and accepted (:212-236), walker 1 the clone
    crowd.mw_recompute()
    lp = crowd.mw_evaluateGL()[0]
    assert lp[0] == pytest.approx(LOGPSI_MOVED, rel=1e-9) and lp[1] == pytest.approx(LOGPSI_0, rel=1e-9)  # :274-277
    g = np.asarray(crowd.mw_evalGrad(0)).reshape(2, 3)  # :289-304
    assert g[0] == pytest.approx(GRAD_OLD[0], rel=1e-7) and g[1] == pytest.approx(GRAD_OLD[1], rel=1e-7)
    crowd.mw_makeMove(0, np.stack([DELTA_SIGN, DELTA_SIGN]))  # :310-329
    ratios, grads = crowd.mw_calcRatioGrad(0)
    ratios, grads = np.asarray(ratios).reshape(2), np.asarray(grads).reshape(2, 3)
    assert ratios[0] == pytest.approx(R_SIGN[0], rel=1e-7) and ratios[1] == pytest.approx(R_SIGN[1], rel=1e-7)
    assert grads[1] == pytest.approx(GRAD_SIGN_1, rel=1e-6)
    crowd.mw_accept_rejectMove(0, [0, 0])  # the reference simply proposes again; here the proposal is rejected first
    crowd.mw_evalGrad(0)
    crowd.mw_makeMove(0, np.stack([np.zeros(3), DELTA]))  # :341-377
    ratios, grads = crowd.mw_calcRatioGrad(0)
    ratios, grads = np.asarray(ratios).reshape(2), np.asarray(grads).reshape(2, 3)
    assert ratios[0] == pytest.approx(1.0, rel=1e-9) and ratios[1] == pytest.approx(RATIO_ALL, rel=1e-8)
    assert grads[0] == pytest.approx(GRAD_OLD[0], rel=1e-7) and grads[1] == pytest.approx(GRAD_OLD[0], rel=1e-7)
    crowd.mw_accept_rejectMove(0, [1, 1])  # :385-399
    crowd.mw_completeUpdates()
    lp = crowd.mw_evaluateGL()[0]
    assert lp[0] == pytest.approx(LOGPSI_MOVED, rel=1e-8) and lp[1] == pytest.approx(LOGPSI_MOVED, rel=1e-8)


# ---------------------------------------------------------------- real orbitals: test_TrialWaveFunction.cpp (1x1x1 cell)
R_PRIM = np.array([[3.37316115, 3.37316115, 0.0], [0.0, 3.37316115, 3.37316115], [3.37316115, 0.0, 3.37316115]])
SHIFTS1 = np.array([np.array(c) @ R_PRIM for c in itertools.product(range(-2, 3), repeat=3)])
WS1 = min(np.linalg.norm(s) for s in SHIFTS1 if np.any(s)) / 2
# literals of the non-complex branch
RL_LOGPSI_0, RL_LOGPSI_MOVED = -1.471840358291562, -0.6365029797784554     # :157, :197
RL_RATIO_ALL, RL_RATIO_FERMI = 2.305591774210242, 2.515045914101833         # :187-188
RL_GRAD_OLD = [[14.77249702264, -20.385235323777, 4.8529516184558], [47.38770710732, -63.361119579044, 15.318325284049]]
RL_R_SIGN = [-0.4138835449, -2.5974770159]                                  # :284-285
RL_GRAD_SIGN_1 = [-17.865723259764, 19.854257889369, -2.9669578650441]      # :286-288


class NumpyTWFReal(NumpyTWF):
    def __init__(self, orc, coefs, G):
        self.orc, self.coefs, self.G = orc, coefs, G

    def spo(self, r):
        psi, dpsi, _ = self.orc.r2r_vgl(self.coefs, self.G, 2, np.atleast_2d(r))
        return psi.astype(np.float64), dpsi.astype(np.float64)

    def u(self, r):
        return [x[0] for x in self.orc.functor_eval(UU, WS1, -0.25, np.array([r]))]

    @staticmethod
    def min_image(dv):
        c = dv + SHIFTS1
        return c[np.argmin((c**2).sum(1))]


def real_table(orc, dtype):
    d = np.load(os.path.join(HERE, "golden", "diamondC_1x1x1_eshdf.npz"))
    return eshdf_spline.build_table(orc, d["psi_g"], d["gvectors"], d["reduced_k"], d["eigenvalues"], 2, dtype)


def test_numpy_restatement_real_branch(orc):
    """precision="float" table, double determinant arithmetic: what the reference's full-precision build runs.  Catch's
    Approx is 1.2e-5 relative; the walker that never moved agrees to 1e-10, the moved one to 1e-7 (FFT library rounding
    in the float table)."""
    w = NumpyTWFReal(orc, real_table(orc, np.float32), np.linalg.inv(R_PRIM))
    l0 = w.logpsi(R0)
    assert l0.real == pytest.approx(RL_LOGPSI_0, rel=1e-6) and abs(l0.imag) == pytest.approx(np.pi)  # :235-237
    R1 = R0.copy()
    R1[0] += DELTA
    l1 = w.logpsi(R1)
    assert l1.real == pytest.approx(RL_LOGPSI_MOVED, rel=1e-6)
    assert np.exp(l1 - l0).real == pytest.approx(RL_RATIO_ALL, rel=1e-6)
    assert np.exp(w.log_det(R1) - w.log_det(R0)).real == pytest.approx(RL_RATIO_FERMI, rel=1e-6)
    assert w.grad(R1, 0) == pytest.approx(RL_GRAD_OLD[0], rel=1e-6)
    assert w.grad(R0, 0) == pytest.approx(RL_GRAD_OLD[1], rel=1e-9)
    R1s, R0s = R1.copy(), R0.copy()
    R1s[0] += DELTA_SIGN
    R0s[0] += DELTA_SIGN
    assert np.exp(w.logpsi(R1s) - l1).real == pytest.approx(RL_R_SIGN[0], rel=1e-6)
    assert np.exp(w.logpsi(R0s) - l0).real == pytest.approx(RL_R_SIGN[1], rel=1e-9)
    assert w.grad(R0s, 0) == pytest.approx(RL_GRAD_SIGN_1, rel=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("dt,rel", [(np.float64, 1.2e-5), (np.float32, 1e-3)])
def test_gpu_twf_sequence_real_branch(orc, dt, rel):
    """full precision (double table; the reference's own Approx tolerance) and mixed precision (float table, float
    inverse; the reference's MIXED_PRECISION tolerances are 1e-3 on gradients and 2e-4 on ratios)"""
    from qmcpack_b200 import api, build
    build.build()
    api.init(0)
    coefs = real_table(orc, dt)
    system = dict(n_up=2, n_dn=2, lattice=R_PRIM, coefs=[coefs, coefs], j2=dict(uu=UU, ud=None, rcut=WS1))
    crowd = api.Crowd(system, nw=2, delay_rank=2)
    R1 = R0.copy()
    R1[0] += DELTA
    crowd.set_positions(np.stack([R1, R0]))
    crowd.mw_recompute()
    lp = crowd.mw_evaluateGL()[0]
    assert lp[0] == pytest.approx(RL_LOGPSI_MOVED, rel=rel) and lp[1] == pytest.approx(RL_LOGPSI_0, rel=rel)
    g = np.asarray(crowd.mw_evalGrad(0)).reshape(2, 3)
    assert g[0] == pytest.approx(RL_GRAD_OLD[0], rel=rel) and g[1] == pytest.approx(RL_GRAD_OLD[1], rel=rel)
    crowd.mw_makeMove(0, np.stack([DELTA_SIGN, DELTA_SIGN]))
    ratios, grads = crowd.mw_calcRatioGrad(0)
    assert ratios[0] == pytest.approx(RL_R_SIGN[0], rel=rel) and ratios[1] == pytest.approx(RL_R_SIGN[1], rel=rel)
    assert np.asarray(grads).reshape(2, 3)[1] == pytest.approx(RL_GRAD_SIGN_1, rel=rel)
    crowd.mw_accept_rejectMove(0, [0, 0])
    crowd.mw_evalGrad(0)
    crowd.mw_makeMove(0, np.stack([np.zeros(3), DELTA]))
    ratios, grads = crowd.mw_calcRatioGrad(0)
    grads = np.asarray(grads).reshape(2, 3)
    assert ratios[0] == pytest.approx(1.0, rel=rel) and ratios[1] == pytest.approx(RL_RATIO_ALL, rel=rel)
    assert grads[0] == pytest.approx(RL_GRAD_OLD[0], rel=rel) and grads[1] == pytest.approx(RL_GRAD_OLD[0], rel=rel)
    crowd.mw_accept_rejectMove(0, [1, 1])
    crowd.mw_completeUpdates()
    lp = crowd.mw_evaluateGL()[0]
    assert lp[0] == pytest.approx(RL_LOGPSI_MOVED, rel=rel) and lp[1] == pytest.approx(RL_LOGPSI_MOVED, rel=rel)


# ---------------------------------------------------------------- the oracle DRIVER itself on the same literals
def test_oracle_driver_on_reference_data(orc, spo_data):
    """OracleVMC (the checker of the GPU VMC tests): its from-scratch evaluation -- spline rows, LU inverse +
    log-determinant, J2 sums -- and the evalGrad / makeMove / calcRatioGrad blocks of its move loop reproduce the reference's
    literals for both cells: complex orbitals in the 2x1x1 tiling (spline on the primitive
    cell) and real orbitals on the 1x1x1 cell."""
    coefs, Gp, kc = spo_data
    d2 = np.load(os.path.join(HERE, "golden", "diamondC_2x1x1_eshdf.npz"))
    sys_c = dict(n_up=2, n_dn=2, lattice=R_SUPER, spline_lattice=d2["primitive_vectors"], coefs=[coefs, coefs],
                 kpts=[kc, kc], j2=dict(uu=UU, ud=None, rcut=WS_RADIUS))
    R1 = R0.copy()
    R1[0] += DELTA
    ov = oracle_lib.OracleVMC(orc, sys_c, nw=2, ncrowds=1, seeds=[1], delay_rank=2)
    ov.set_positions(np.stack([R1, R0]))
    ov.recompute()
    lp = ov.evaluate_gl()[0]
    assert lp[0] == pytest.approx(LOGPSI_MOVED, rel=1e-9) and lp[1] == pytest.approx(LOGPSI_0, rel=1e-9)
    # the move path of the driver (evalGrad / makeMove / calcRatioGrad blocks of advanceCrowd) on the prescribed moves
    r0, go0, _ = ov.probe_move(0, 0, DELTA_SIGN)
    r1, go1, gn1 = ov.probe_move(1, 0, DELTA_SIGN)
    assert go0 == pytest.approx(GRAD_OLD[0], rel=1e-8) and go1 == pytest.approx(GRAD_OLD[1], rel=1e-8)
    assert r0 == pytest.approx(R_SIGN[0], rel=1e-8) and r1 == pytest.approx(R_SIGN[1], rel=1e-8)
    assert gn1 == pytest.approx(GRAD_SIGN_1, rel=1e-7)
    r1, _, gn1 = ov.probe_move(1, 0, DELTA)
    assert r1 == pytest.approx(RATIO_ALL, rel=1e-9) and gn1 == pytest.approx(GRAD_OLD[0], rel=1e-8)
    r0, _, gn0 = ov.probe_move(0, 0, np.zeros(3))
    assert r0 == pytest.approx(1.0, rel=1e-12) and gn0 == pytest.approx(GRAD_OLD[0], rel=1e-8)
    # second half of the reference test (:426-505): both walkers at the accepted position, electron 1 next
    DISPL_NEXT = np.array([0.1, 0.2, 0.3])
    ov.set_positions(np.stack([R1, R1]))
    ov.recompute()
    grad_next = [complex(-114.82740072726, -7.605305979232e-05), complex(-93.980772428401, -7.605302517238e-05),
                 complex(64.050803536571, 7.6052975324197e-05)]          # :436-447
    grad_next_new = [complex(9.6073058494562, -1.4375146770852e-05), complex(6.3111018321898, -1.4375146510386e-05),
                     complex(-3.2027658046121, 1.4375146020225e-05)]     # :466-477
    for iw in (0, 1):
        _, go, gn = ov.probe_move(iw, 1, DISPL_NEXT)
        assert go == pytest.approx(grad_next, rel=1e-7) and gn == pytest.approx(grad_next_new, rel=1e-7)
    # walker 0 accepts that move; the inverse of its up determinant after the update (:499-502, psiMinv = (psiM^-1)^T)
    R2 = R1.copy()
    R2[1] += DISPL_NEXT
    ov.set_positions(np.stack([R2, R1]))
    ov.recompute()
    minv = ov.psiminv(0, 0)[0]
    want = np.array([[complex(38.503358805635, -38.503358805645), complex(-31.465077529568, 31.465077529576)],
                     [complex(-27.188228530061, 27.188228530068), complex(22.759962501254, -22.75996250126)]])
    assert minv == pytest.approx(want, rel=1e-7)
    # (the reference's literals come from a float table under double determinants; the oracle driver is all-double or
    # all-float, so the double table is checked within the reference's own Approx and the float one within 1e-4)
    for dt, rel in ((np.float64, 1.2e-5), (np.float32, 1e-4)):
        tab = real_table(orc, dt)
        sys_r = dict(n_up=2, n_dn=2, lattice=R_PRIM, coefs=[tab, tab], j2=dict(uu=UU, ud=None, rcut=WS1))
        ov = oracle_lib.OracleVMC(orc, sys_r, nw=2, ncrowds=1, seeds=[1], delay_rank=2)
        ov.set_positions(np.stack([R1, R0]))
        ov.recompute()
        lp = ov.evaluate_gl()[0]
        assert lp[0] == pytest.approx(RL_LOGPSI_MOVED, rel=rel) and lp[1] == pytest.approx(RL_LOGPSI_0, rel=rel)
        r0, go0, _ = ov.probe_move(0, 0, DELTA_SIGN)
        r1, go1, gn1 = ov.probe_move(1, 0, DELTA_SIGN)
        assert go0.real == pytest.approx(RL_GRAD_OLD[0], rel=rel) and go1.real == pytest.approx(RL_GRAD_OLD[1], rel=rel)
        assert r0.real == pytest.approx(RL_R_SIGN[0], rel=rel) and r1.real == pytest.approx(RL_R_SIGN[1], rel=rel)
        assert gn1.real == pytest.approx(RL_GRAD_SIGN_1, rel=rel)
        r1, _, gn1 = ov.probe_move(1, 0, DELTA)
        assert r1.real == pytest.approx(RL_RATIO_ALL, rel=rel) and gn1.real == pytest.approx(RL_GRAD_OLD[0], rel=rel)
