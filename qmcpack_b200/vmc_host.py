"""Host-side restatement of VMCBatched::advanceWalkers' particle-by-particle loop
(reference: src/QMCDrivers/VMC/VMCBatched.cpp:106-176) driving the C ABI one mw_* call at a time, exactly the call
sequence the reference's TWFdispatcher / PSdispatcher issue:

    flex_evalGrad -> flex_makeMove -> flex_calcRatioGrad -> (accept test on the host) -> flex_accept_rejectMove

The random number generator is injected (`rng.gauss(n, dtype)`, `rng.uniform()`): tests pass the oracle's StdRandom so
that the product and the CPU reference consume one and the same std::mt19937 stream.  Used by tests and as the
readable specification of csrc/vmc_host.cpp (the compiled driver bench.py times for the e2e number).
"""
import numpy as np


def get_drift(tau, g, RT):
    """DriftModifierUNR::getDrift, a = 1 (GreenFunctionModifiers/DriftModifierUNR.cpp:20-31); g is [nw][3] in RT."""
    g = g.astype(RT)
    vsq = (g[:, 0] * g[:, 0] + g[:, 1] * g[:, 1] + g[:, 2] * g[:, 2]).astype(RT)
    eps = np.finfo(RT).eps
    with np.errstate(divide="ignore", invalid="ignore"):
        sc_big = ((-1.0 + np.sqrt(1.0 + 2.0 * np.float64(RT(1)) * np.float64(tau) * vsq.astype(np.float64))) /
                  (RT(1) * vsq).astype(np.float64)).astype(RT)
    sc = np.where(vsq < eps, RT(tau), sc_big).astype(RT)
    return (g * sc[:, None]).astype(RT)


def advance_walkers(crowd, rng, tau=0.3, use_drift=True, log_accept=None, log_ratio=None):
    """One sweep (sub_steps = 1) over all electrons of all walkers of `crowd`; returns the number of accepted moves."""
    RT = crowd.T
    nw, N = crowd.nw, crowd.N
    tauovermass = RT(tau) * RT(1.0)
    oneover2tau = RT(0.5 / tauovermass)
    sqrttau = RT(np.sqrt(tauovermass))
    walker_deltas = rng.gauss(3 * nw * N, RT).reshape(N, nw, 3)  # element iat*nw + iw (VMCBatched.cpp:122)
    eps = np.finfo(RT).eps
    n_acc = 0
    for iat in range(N):
        deltas = (walker_deltas[iat] * sqrttau).astype(RT)
        if use_drift:
            grads_now = np.real(crowd.mw_evalGrad(iat)).astype(RT)  # convertToReal of a complex gradient
            drifts = (get_drift(tauovermass, grads_now, RT) + deltas).astype(RT)
        else:
            drifts = deltas
        crowd.mw_makeMove(iat, drifts.astype(np.float64))
        ratios, grads_new = crowd.mw_calcRatioGrad(iat)
        log_gf = np.zeros(nw, RT)
        log_gb = np.zeros(nw, RT)
        if use_drift:
            log_gf = (-oneover2tau * (deltas * deltas).sum(axis=1, dtype=RT)).astype(RT)
            rev = (get_drift(tauovermass, np.real(grads_new).astype(RT), RT) + drifts).astype(RT)
            log_gb = (-oneover2tau * (rev * rev).sum(axis=1, dtype=RT)).astype(RT)
        prob = (np.abs(ratios) ** 2 if np.iscomplexobj(ratios) else ratios * ratios).astype(RT)  # std::norm
        accepted = np.zeros(nw, np.uint8)
        for iw in range(nw):  # the uniform is drawn only when prob >= eps (VMCBatched.cpp:156-158)
            if prob[iw] >= eps and rng.uniform() < np.float64(RT(prob[iw] * np.exp(RT(log_gb[iw] - log_gf[iw])))):
                accepted[iw] = 1
        crowd.mw_accept_rejectMove(iat, accepted, True)
        n_acc += int(accepted.sum())
        if log_accept is not None:
            log_accept[iat] = accepted
        if log_ratio is not None:
            log_ratio[iat] = ratios
    crowd.mw_completeUpdates()
    return n_acc
