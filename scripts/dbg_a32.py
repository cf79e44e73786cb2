import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, oracle_lib
from qmcpack_b200 import workload, api, vmc_host
api.init(0)
orc=oracle_lib.port()
for dt in (np.float32, np.float64):
    s=workload.make_system(N=384, M=48, dtype=np.float32)
    if dt==np.float64:
        s["coefs"]=[workload.aligned_zeros(c.shape, np.float64) for c in s["coefs"]]
        s0=workload.make_system(N=384, M=48, dtype=np.float32)
        for a,b in zip(s["coefs"], s0["coefs"]): a[...]=b
    nw,seed,tau,N=4,31,0.3,384
    R=workload.initial_positions(s,nw)
    crowd=api.Crowd(s,nw=nw,delay_rank=32); crowd.set_positions(R); crowd.mw_recompute()
    rng=orc.rng(seed)
    log=np.zeros((1,N,nw),np.uint8); ratios=np.zeros((1,N,nw))
    vmc_host.advance_walkers(crowd,rng,tau=tau,log_accept=log[0],log_ratio=ratios[0])
    ov=oracle_lib.OracleVMC(orc,s,nw=nw,ncrowds=1,seeds=[seed],tau=tau,delay_rank=32)
    ov.set_positions(R); ov.recompute()
    orat=ov.sweep_forced(log)
    rel=np.abs(ratios-orat)/np.maximum(np.abs(orat),1e-3)
    print(dt.__name__, "acc", log.mean())
    for c in range(0,N,32):
        print("  iat %3d-%3d max rel per walker"%(c,c+31), rel[0,c:c+32].max(axis=0), "min|ratio|", np.abs(orat[0,c:c+32]).min(axis=0))
