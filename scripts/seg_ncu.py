"""one warm NiO-a64 sweep on the walker-segment kernel, bracketed by cudaProfilerStart/Stop (ncu --profile-from-start off)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qmcpack_b200 import api, workload
api.init(0)
cfg = os.environ.get("SEG_CONFIG", "NiO-a64")
nw = int(os.environ.get("SEG_NW", "512"))
c = workload.CONFIGS[cfg]
s = workload.make_system(N=c["N"], M=c["M"], dtype=c["dtype"])
crowd = api.Crowd(s, nw=nw, delay_rank=c["k"])
crowd.set_positions(workload.initial_positions(s, nw))
crowd.mw_recompute()
crowd.vmc_init(tau=0.3, use_drift=True, seed=1000, use_cuda_graph=False, sweep_kernel=int(os.environ.get("SEG_SK", "2")))
crowd.vmc_sweep(1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
crowd.vmc_sweep(1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", crowd.sweep_kernel)
