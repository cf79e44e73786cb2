import sys, os, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import oracle_lib
from qmcpack_b200 import api, build, vmc_host
from qmcpack_b200.workload import make_system, initial_positions
build.build(); api.init(0)
orc = oracle_lib.port()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dt = np.float64
nw, tau, seed = 2, 0.3, 77
s = make_system(N=N, M=20, dtype=dt)
R = initial_positions(s, nw)
crowd = api.Crowd(s, nw=nw, delay_rank=k)
crowd.set_positions(R); crowd.mw_recompute()
rng = orc.rng(seed)
log = np.zeros((1, N, nw), np.uint8); ratios = np.zeros((1, N, nw))
vmc_host.advance_walkers(crowd, rng, tau=tau, log_accept=log[0], log_ratio=ratios[0])
ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[seed], tau=tau, delay_rank=k)
ov.set_positions(R); ov.recompute()
oratios = ov.sweep_forced(log)
rel = np.abs(ratios - oratios) / np.maximum(np.abs(oratios), 0.1)
bad = np.argwhere(rel[0] > 1e-7)
print("N", N, "k", k, "first bad", bad[:5].tolist(), "max", rel.max())
for i in range(0, N, max(1, N // 48)):
    print(i, rel[0, i].max())
