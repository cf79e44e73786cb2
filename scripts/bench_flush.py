"""Times the rank-k Woodbury flush (DelayedUpdateBatched::mw_updateInvMat) alone through the measurement hook of the C ABI
(qmcb_det_time_update_inv_mat: CUDA events on the crowd stream, back-to-back launches) and prints one JSON line per case:
executed flop rate against the measured FP64 tensor peak (scripts/micro/mma_rate.cu: DMMA 37.0 TF/s on B200) and one-pass
bytes against the HBM peak.  QMCB_FLUSH=simt selects the three-GEMM SIMT form for comparison (read once per process).

  python scripts/bench_flush.py --dtype f64 --n 384 --k 32 --walkers 512
  python scripts/bench_flush.py --dtype c128 --n 768 --k 64 --walkers 128
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="f64", choices=["f32", "f64", "c64", "c128"])
ap.add_argument("--n", type=int, default=384)
ap.add_argument("--k", type=int, default=32)
ap.add_argument("--walkers", type=int, default=512)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()

from qmcpack_b200 import api
from qmcpack_b200.workload import random_table

dt = {"f32": np.float32, "f64": np.float64, "c64": np.complex64, "c128": np.complex128}[args.dtype]
dt = np.dtype(dt)
n = args.n
api.init(0)
if dt.kind == "c":
    t = random_table((4, 4, 4), 2 * n, np.float32 if dt == np.complex64 else np.float64, seed=1)
    kp = np.tile([0.1, 0.2, 0.3], (n, 1))
    system = dict(n_up=n, n_dn=n, lattice=np.eye(3) * 4.0, coefs=[t, t], kpts=[kp, kp])
else:
    t = random_table((4, 4, 4), n, dt, seed=1)
    system = dict(n_up=n, n_dn=n, lattice=np.eye(3) * 4.0, coefs=[t, t])
crowd = api.Crowd(system, nw=args.walkers, delay_rank=args.k)
us = crowd.det_time_update_inv_mat(0, args.k, args.reps)
cplx = 4 if dt.kind == "c" else 1
flops = cplx * (4.0 * args.k * n * n + 2.0 * n * args.k * args.k) * args.walkers  # 2kn^2 + 2nk^2 + 2n^2k
bytes_one_pass = 2.0 * n * n * dt.itemsize * args.walkers
print(json.dumps({
    "flush": os.environ.get("QMCB_FLUSH", "default"), "dtype": args.dtype, "n": n, "k": args.k, "walkers": args.walkers,
    "us_per_flush": us, "tflops_executed": flops / us * 1e-6, "frac_fp64_dmma_peak_37.0": flops / us * 1e-6 / 37.0,
    "one_pass_GBps": bytes_one_pass / us * 1e-3, "dmma_split": os.environ.get("QMCB_DMMA_SPLIT", "ntiles")}))
