"""Spline gather kernel alone at a named NiO shape: CUDA-event time per launch, algorithmic GB/s vs measured HBM peak."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qmcpack_b200 import api, workload

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="NiO-a64")
ap.add_argument("--walkers", type=int, default=512)
ap.add_argument("--reps", type=int, default=40)
ap.add_argument("--iid", type=int, default=1)
args = ap.parse_args()
api.init(0)
c = workload.CONFIGS[args.config]
n = c["N"] // 2
T = c["dtype"]
L = workload.L_A64 * (c["N"] / 768.0) ** (1 / 3)
lat = np.eye(3) * L
coefs = workload.random_table(c["M"], n, T, seed=1)
spo = api.SplineSPOSet(coefs, n, np.linalg.inv(lat))
nw = args.walkers
tdt = torch.float32 if T == np.float32 else torch.float64
gen = torch.Generator(device="cuda").manual_seed(5)
nsets = 24
pos = (torch.rand((nsets, nw, 3), generator=gen, device="cuda", dtype=torch.float64) * L).to(tdt).contiguous()
inv = torch.randn((nw, n), generator=gen, device="cuda", dtype=tdt).contiguous()
phi = torch.empty((5, nw, n), device="cuda", dtype=tdt)
rg = torch.empty((nw, spo.rg_parts, 4), device="cuda", dtype=tdt)
ts = torch.cuda.Stream()
lib = api.lib()
def launch(i):
    rc = lib.qmcb_spline_mw_vgl_ratio_grads_dev(spo.h, nw, pos[i % nsets].data_ptr(), inv.data_ptr(), n, phi.data_ptr(), rg.data_ptr(), ts.cuda_stream)
    assert rc == 0, lib.qmcb_last_error()
for i in range(5): launch(i)
ts.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(ts)
for i in range(args.reps): launch(5 + i)
e1.record(ts); ts.synchronize()
t = e0.elapsed_time(e1) * 1e-3 / args.reps
npad = workload.aligned_size(T, n); es = np.dtype(T).itemsize
b = (64 * npad + 6 * n) * es * nw
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
print(json.dumps({"config": args.config, "walkers": nw, "us_per_launch": t * 1e6, "GBps": b / t / 1e9, "frac": b / t / 1e9 / pk, "evals_per_s": nw / t}))
