"""first-move comparison product vs oracle on the non-reduced (sheared) cell"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from qmcpack_b200 import api, workload
import oracle_lib
api.init(0)
orc = oracle_lib.port()
LAT_GENERAL = np.array([[6.0, 0.4, 0.0], [0.3, 6.5, -0.2], [0.1, -0.3, 7.0]])
LAT_SHEARED = np.array([[1, 0, 0], [2, 1, 0], [1, -1, 1]], float) @ LAT_GENERAL
for name, lat in (("general", LAT_GENERAL), ("sheared", LAT_SHEARED)):
    for wj in ((True, True), (False, False), (True, False), (False, True)):
        s = workload.make_system(N=24, M=8, dtype=np.float64, L=6.0, lattice=lat, with_j1=wj[0], with_j2=wj[1])
        nw = 3
        R = workload.initial_positions(s, nw)
        crowd = api.Crowd(s, nw=nw, delay_rank=4)
        crowd.set_positions(R); crowd.mw_recompute()
        ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[1], tau=0.1, delay_rank=4)
        ov.set_positions(R); ov.recompute()
        lp, ke, _, _ = crowd.mw_evaluateGL(); olp, oke, _, _ = ov.evaluate_gl()
        displ = np.random.default_rng(3).normal(size=(nw, 3)) * 0.3
        g = crowd.mw_evalGrad(0)
        crowd.mw_makeMove(0, displ)
        r, gn = crowd.mw_calcRatioGrad(0)
        for iw in range(nw):
            orr, ogo, ogn = ov.probe_move(iw, 0, displ[iw])
            print(name, "j1,j2", wj, "iw", iw, "logpsi", lp[iw] - olp[iw], "ke", ke[iw] - oke[iw], "ratio", r[iw], orr.real,
                  "grad_old", np.abs(g[iw] - ogo.real).max(), "grad_new", np.abs(gn[iw] - ogn.real).max())
        crowd.mw_accept_rejectMove(0, np.zeros(nw, np.uint8), True)
