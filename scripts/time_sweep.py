"""In-situ cost attribution: time graph-replayed sweeps with parts of the system switched off.
python scripts/time_sweep.py [--walkers 512] [--crowds 1]"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="NiO-a64")
ap.add_argument("--walkers", type=int, default=512)
ap.add_argument("--crowds", type=int, default=1)
ap.add_argument("--steps", type=int, default=4)
args = ap.parse_args()
import torch
from qmcpack_b200 import api, workload
api.init(0)
c = workload.CONFIGS[args.config]
base = workload.make_system(N=c["N"], M=c["M"], dtype=c["dtype"])
spo = None
for name, drop in (("full", ()), ("no J1", ("j1",)), ("no J1, no J2", ("j1", "j2"))):
    s = {k: v for k, v in base.items() if k not in drop}
    nwc = args.walkers // args.crowds
    crowds = []
    for i in range(args.crowds):
        cr = api.Crowd(s, nw=nwc, delay_rank=c["k"], spo=spo)
        spo = cr.spo
        cr.set_positions(workload.initial_positions(s, nwc, seed=7 + 1000 * i))
        cr.mw_recompute()
        cr.vmc_init(tau=0.3, use_drift=True, seed=1000 + i, use_cuda_graph=True)
        crowds.append(cr)
    for _ in range(3):
        for cr in crowds:
            cr.vmc_sweep_async()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for cr in crowds:
            cr.vmc_sweep_async()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    print(f"{name:16s} crowds {args.crowds}: {dt * 1e3:.2f} ms/sweep = {dt * 1e6 / c['N']:.1f} us/move, "
          f"{args.walkers * c['N'] / dt / 1e6:.2f} M moves/s")
    del crowds
