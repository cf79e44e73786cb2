"""launch list of one sweep WITHOUT Jastrow factors (cost attribution); use under ncu like profile_sweep.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qmcpack_b200 import api, workload
api.init(0)
c = workload.CONFIGS["NiO-a64"]
s = workload.make_system(N=c["N"], M=c["M"], dtype=c["dtype"], with_j1=False, with_j2=False)
crowd = api.Crowd(s, nw=512, delay_rank=c["k"])
crowd.set_positions(workload.initial_positions(s, 512))
crowd.mw_recompute()
crowd.vmc_init(tau=0.3, use_drift=True, seed=1000, use_cuda_graph=False)
crowd.vmc_sweep(1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
crowd.vmc_sweep(1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
