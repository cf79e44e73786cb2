"""tiny run of the walker-segment kernel (for compute-sanitizer / cuda-gdb on the GPU box)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from qmcpack_b200 import api, workload
api.init(0)
dt = np.float64 if "f64" in sys.argv else np.float32
N = int(os.environ.get("SEG_N", "24"))
s = workload.make_system(N=N, M=8, dtype=dt, L=6.0 if N <= 48 else None)
nw, k = int(os.environ.get("SEG_NW", "5")), int(os.environ.get("SEG_K", "4"))
crowd = api.Crowd(s, nw=nw, delay_rank=min(k, N // 2))
crowd.set_positions(workload.initial_positions(s, nw))
crowd.mw_recompute()
crowd.vmc_init(tau=0.1, use_drift=True, seed=5, use_cuda_graph=False, sweep_kernel=2)
log = crowd.vmc_sweep(2, log_accept=True)
print("ok", log.mean(), crowd.sweep_kernel)
if "check" in sys.argv:
    import oracle_lib
    orc = oracle_lib.port()
    ov = oracle_lib.OracleVMC(orc, s, nw=nw, ncrowds=1, seeds=[5], tau=0.1, delay_rank=min(k, N // 2))
    ov.set_positions(workload.initial_positions(s, nw))
    ov.recompute()
    olog = ov.sweep(2, log_accept=True)
    print("identical acceptance:", np.array_equal(log, olog), "first diffs", np.argwhere(log != olog)[:5].tolist())
