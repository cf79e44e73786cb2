"""Host-driven path alone (the e2e leg of bench.py): compiled host VMC driver above the C ABI, one host thread per crowd.
python scripts/time_e2e.py [--walkers 512] [--crowds 4] [--steps 3]"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="NiO-a64")
ap.add_argument("--walkers", type=int, default=512)
ap.add_argument("--crowds", type=int, default=4)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
import torch
from qmcpack_b200 import api, workload
api.init(0)
c = workload.CONFIGS[args.config]
s = workload.make_system(N=c["N"], M=c["M"], dtype=c["dtype"])
nw, ncr = args.walkers, args.crowds
R = workload.initial_positions(s, nw)
base, extra = divmod(nw, ncr)
sizes = [base + (1 if i < extra else 0) for i in range(ncr)]
crowds, off, spo = [], 0, None
for i in range(ncr):
    cr = api.Crowd(s, nw=sizes[i], delay_rank=c["k"], spo=spo)
    spo = cr.spo
    cr.set_positions(R[off:off + sizes[i]])
    cr.mw_recompute()
    crowds.append(cr)
    off += sizes[i]
drv = api.HostVMC(crowds, [2000 + i for i in range(ncr)], tau=0.3, use_drift=True)
drv.run(1)
torch.cuda.synchronize()
t0 = time.perf_counter()
drv.run(args.steps)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / args.steps
a, r = drv.counts()
print(f"e2e host-driven: crowds {ncr}: {dt * 1e3:.2f} ms/sweep = {dt * 1e6 / c['N']:.1f} us/move, "
      f"{nw * c['N'] / dt / 1e6:.2f} M moves/s, acceptance {a / (a + r):.3f}, "
      f"env HOST_FUSE={os.environ.get('QMCB_HOST_FUSE', 'default')}")
