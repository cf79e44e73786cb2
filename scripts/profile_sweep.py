"""Profiling harness: sets the NiO-a64 crowd up, runs warm sweeps, then brackets ONE sweep (no CUDA graph, so every kernel
is a plain launch) with cudaProfilerStart/Stop.  Use under ncu with --profile-from-start off:

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv \
      --log-file gpurun_out/launches.csv python scripts/profile_sweep.py
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:spline_gather -c 3 \
      -o gpurun_out/spline python scripts/profile_sweep.py
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="NiO-a64")
ap.add_argument("--walkers", type=int, default=512)
ap.add_argument("--graph", type=int, default=0)
args = ap.parse_args()

import torch
from qmcpack_b200 import api, workload

api.init(0)
c = workload.CONFIGS[args.config]
s = workload.make_system(N=c["N"], M=c["M"], dtype=c["dtype"])
crowd = api.Crowd(s, nw=args.walkers, delay_rank=c["k"])
crowd.set_positions(workload.initial_positions(s, args.walkers))
crowd.mw_recompute()
crowd.vmc_init(tau=0.3, use_drift=True, seed=1000, use_cuda_graph=bool(args.graph))
crowd.vmc_sweep(1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
crowd.vmc_sweep(1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one sweep; acceptance", crowd.vmc_counts()[0].sum() / (2 * args.walkers * c["N"]))
