"""mw_recompute's FP64 inverse + log-determinant at a benchmark shape: own blocked Gauss-Jordan kernels vs cuBLAS
getrf/getriBatched (qmcb_det_time_inverse).  python scripts/time_inverse.py [--config NiO-a64] [--walkers 512] [--reps 3]
Under ncu (--metrics gpu__time_duration.sum) the launch list shows the panel / update split."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="NiO-a64")
ap.add_argument("--walkers", type=int, default=512)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--methods", default="2,1")
args = ap.parse_args()
from qmcpack_b200 import api, workload
api.init(0)
c = workload.CONFIGS[args.config]
cplx = bool(c.get("complex_orbitals"))
n = c["N"] // 2
# only the determinant engine is exercised: a tiny table, the benchmarked matrix size
t = workload.pw_table((12, 12, 12), n, np.float64, 3) if not cplx else workload.pw_table_complex((12, 12, 12), n, np.float64, 3)
s = dict(n_up=n, n_dn=n, lattice=np.eye(3) * 6.0 * (n / 12) ** (1 / 3), coefs=[t, t])
if cplx:
    kp = np.tile([0.1, 0.2, 0.3], (n, 1))
    s["kpts"] = [kp, kp]
crowd = api.Crowd(s, nw=args.walkers, delay_rank=c["k"])
crowd.set_positions(workload.initial_positions(s, args.walkers))
crowd.mw_recompute()
out = dict(config=args.config, walkers=args.walkers, n=n, complex=cplx)
fl = (4 if cplx else 1) * 2.0 * n ** 3 * args.walkers
for m in [int(x) for x in args.methods.split(",")]:
    us = crowd.det_time_inverse(0, m, reps=args.reps)
    out["us_own" if m == 2 else "us_cublas"] = us
    out["tflops_own" if m == 2 else "tflops_cublas"] = fl / us * 1e-6
print(json.dumps(out))
