"""Per-kernel attribution of one NiO-a64 sweep (qmcb_vmc_profile_sweep): walker-segment kernel vs two-kernel path, with and
without the Jastrow factors, at 256 and 512 walkers per crowd.  python scripts/seg_profile.py [--config NiO-a64]"""
import argparse, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="NiO-a64")
ap.add_argument("--walkers", default="256,512")
ap.add_argument("--nojas", type=int, default=1)
args = ap.parse_args()
from qmcpack_b200 import api, workload
api.init(0)
c = workload.CONFIGS[args.config]
variants = [("J1+J2", dict())] + ([("no Jastrow", dict(with_j1=False, with_j2=False))] if args.nojas else [])
for name, kw in variants:
    s = workload.make_system(N=c["N"], M=c["M"], dtype=c["dtype"], **kw)
    spo = None
    for nw in [int(x) for x in args.walkers.split(",")]:
        for sk in (2, 1):
            crowd = api.Crowd(s, nw=nw, delay_rank=c["k"], spo=spo)
            spo = crowd.spo
            crowd.set_positions(workload.initial_positions(s, nw))
            crowd.mw_recompute()
            try:
                crowd.vmc_init(tau=0.3, use_drift=True, seed=1000, use_cuda_graph=False, sweep_kernel=sk)
            except RuntimeError as e:
                print(name, nw, sk, "unavailable:", e)
                continue
            crowd.vmc_sweep(1)
            p = crowd.vmc_profile_sweep()
            p = crowd.vmc_profile_sweep()
            moves = c["N"]
            print(json.dumps(dict(variant=name, walkers=nw, sweep_kernel=sk, us_per_move=p["sweep_us"] / moves,
                                  Mmoves_per_s=nw * moves / p["sweep_us"], **{k: round(v, 1) for k, v in p.items()})))
            del crowd
