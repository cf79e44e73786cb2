#!/usr/bin/env python
"""bench.py -- NiO-a64 batched VMC electron-moves/s (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W]            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                            the reference's CPU implementation of the same path

One "step" = one particle-by-particle VMC sweep (every electron of every walker proposed once: drift + Gaussian
move, distance rows, spline SPO evaluation, determinant ratio/gradient, J1/J2, Metropolis test, delayed update).

  value     whole-job electron-moves/s with everything resident in HBM: the device-resident sweep (on-device
            std::mt19937 + Metropolis test, one CUDA graph per sweep; the persistent walker-segment kernel where the
            wavefunction is eligible), timed with CUDA events on the launching stream, max over ranks.  The timed region
            is one VMC BLOCK: K sweeps, then the block estimator (kinetic energy per walker) and its all-reduce over
            ranks (the path's only collective, EstimatorManagerNew.cpp:338,363).
  e2e       the same metric through the reference-facing C ABI with HOST buffers: the compiled host driver
            (include/qmcb_driver.h) issues evalGrad / makeMove / calcRatioGrad / accept_reject per electron, positions,
            gradients, ratios and accept flags cross PCIe every move, accept test on the host -- how QMCPACK's batched
            driver would call this library.  One host thread per crowd (--crowds, default 8, never more than this rank's
            share of the cores); `host_kernel_mode` 2 = the calls were mailbox exchanges with a resident walker-segment
            kernel, 1 = kernel launches per call.
  roofline  the dominant kernel of the sweep.  With the walker-segment kernel: its launches timed with CUDA events in
            one profiled sweep (qmcb_vmc_profile_sweep), algorithmic bytes per launch = SURVEY 8d's per-move figure
            (64*Npad*4 stencil + 5*n*4 + n*4) x walkers x moves per launch, against the measured HBM copy bandwidth of
            MEASURED_PEAKS.json.  `spline_gather` carries the stand-alone gather kernel (SPOSet::mw_evaluateVGLandDetRatioGrads)
            at the full and at the per-crowd launch size.
  flush     the rank-k Woodbury flush (mw_updateInvMat) timed alone through the qmcb_det_time_update_inv_mat hook:
            algorithmic TF/s, one-pass GB/s; `flush_fp64` the same for the full-precision shapes (NiO-a64 real k = 32,
            NiO-a128 complex k = 64) against `fp64_peak` = cublasDgemm 8192^3 measured in this run.
  recompute the FP64 inverse + log-determinant of mw_recompute on the benchmark crowd's own matrices: this library's blocked
            Gauss-Jordan kernels beside cublas<t>getrfBatched + getriBatched, the routines the reference calls.
  dmc       three DMC generations (device sweep with the DMC rule + C++ WalkerControl::branch) with the all-reduce and the
            packed-walker ncclSend/Recv inside the timed region; --dmc makes it the headline.
  cpu_baseline  the reference's CPU path (oracle/_ref: spline2::evaluate_vgh_impl + DelayedUpdate<T> + DiracMatrix compiled
            from /root/reference, else the oracle port) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "NiO-a64 VMC electron-moves/s"
UNIT = "electron-moves/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="NiO-a64")
    ap.add_argument("--walkers", type=int, default=512, help="walkers per GPU")
    ap.add_argument("--crowds", type=int, default=8, help="crowds (host threads / streams) of the e2e host driver")
    ap.add_argument("--device-crowds", type=int, default=1,
                    help="crowds of the device-resident sweep: walkers/GPU are split over this many crowds, each with its own "
                         "RNG stream and CUDA graph on its own stream (QMCDriverNew gives every crowd its own generator)")
    ap.add_argument("--tau", type=float, default=0.3)
    ap.add_argument("--cpu-walkers", type=int, default=0, help="walkers of the CPU sample (default 2 per core)")
    ap.add_argument("--sweep-kernel", type=int, default=0, help="0 automatic, 1 two-kernel path, 2 walker-segment kernel")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fp64", action="store_true", help="skip the FP64 peak / full-precision flush objects")
    ap.add_argument("--no-recompute", action="store_true", help="skip the mw_recompute inverse object (own kernels vs cuBLAS)")
    ap.add_argument("--dmc", action="store_true",
                    help="headline = batched DMC generations (BASELINE config 5: device move loop + block of C++ WalkerControl::"
                         "branch with the NCCL all-reduce and the packed-walker exchange inside the timed region)")
    ap.add_argument("--no-dmc", action="store_true", help="skip the short DMC object of the default (VMC) line")
    ap.add_argument("--dmc-tau", type=float, default=0.002)
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def workload_desc(cfg, args):
    from qmcpack_b200 import workload
    c = workload.CONFIGS[cfg]
    n = c["N"] // 2
    cplx = bool(c.get("complex_orbitals"))
    npad = workload.aligned_size(c["dtype"], n * (2 if cplx else 1))
    tab_mb = (c["M"] + 3) ** 3 * npad * np.dtype(c["dtype"]).itemsize / 1e6
    return (f"{cfg} synthetic: {c['N']} electrons, {n} {'complex (SplineC2C) ' if cplx else ''}orbitals/spin, {c['M']}^3 spline grid "
            f"({tab_mb:.0f} MB/spin, {np.dtype(c['dtype']).name}), J1+J2 B-spline Jastrows, batched VMC with drift "
            f"(tau={args.tau}), delay_rank {c['k']}, {args.walkers} walkers/GPU")


def ncu_traffic(kernel_substr="spline_gather_kernel"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed ncu --set full
    summary (profiles/), or None"""
    import csv
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_full_*_summary.csv"))):
        try:
            rows = list(csv.reader(open(f)))
            hdr, units = rows[0], rows[1]
            i_name, i_r, i_w = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            vals = [float(r[i_r].replace(",", "")) * scale.get(units[i_r], 1.0) +
                    float(r[i_w].replace(",", "")) * scale.get(units[i_w], 1.0)
                    for r in rows[2:] if kernel_substr in r[i_name]]
            if vals:
                best = {"bytes_per_launch": sum(vals) / len(vals), "source": os.path.relpath(f, ROOT)}
        except Exception:
            pass
    return best


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler:
    """samples SM clock and throttle reasons during the timed region (NVML)"""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            pass
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.ok:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_run(cfg, args, steps, warmup, nw_cpu=None):
    """the reference's CPU implementation of the path on the host cores (bounded sample of the workload)"""
    import oracle_lib
    from qmcpack_b200 import workload
    orc = oracle_lib.ref() or oracle_lib.port()
    kind = "reference" if orc.is_reference else "port"
    cores = os.cpu_count() or 1
    c = workload.CONFIGS[cfg]
    nw = nw_cpu or args.cpu_walkers or 8 * cores
    ncrowds = min(cores, nw)
    orc.lib.orc_set_threads(ncrowds)  # torchrun exports OMP_NUM_THREADS=1: the CPU arm must use every host core
    s = workload.make_system(N=c["N"], M=c["M"], dtype=c["dtype"], complex_orbitals=bool(c.get("complex_orbitals")))
    v = orc.vmc(s, nw=nw, ncrowds=ncrowds, seeds=[1000 + i for i in range(ncrowds)], tau=args.tau, use_drift=True,
                delay_rank=c["k"], batched_engine=False)
    v.set_positions(workload.initial_positions(s, nw))
    v.recompute()
    if warmup:
        v.sweep(warmup)
    t0 = time.perf_counter()
    v.sweep(steps)
    dt = time.perf_counter() - t0
    moves = steps * nw * c["N"]
    acc, rej = v.counts()
    return dict(value=moves / dt, seconds=dt, kind=kind, cores=ncrowds, nw=nw, steps=steps,
                acceptance=float(acc.sum() / max(1, (acc + rej).sum())),
                sample=f"{steps} sweeps of {nw} walkers ({moves} moves) of {cfg}, {ncrowds} crowds = {ncrowds} OpenMP "
                       f"threads, {'oracle/_ref (reference kernels + OpenBLAS)' if kind == 'reference' else 'oracle port'}")


def run_reference(args, rank, world):
    if rank != 0:
        return
    res = cpu_reference_run(args.config, args, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC.replace("NiO-a64", args.config), "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc(args.config, args), "cpu_sample": res["sample"]},
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": res["kind"],
                         "sample": res["sample"], "acceptance": res["acceptance"]},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
def run_dmc_generations(api, torch, dist, s, spo, k, N, nw, rank, local_rank, world, tau, generations, warm):
    """Batched DMC on one crowd per rank: qmcb_vmc_sweep(dmc = 1) + the C++ DMC layer (csrc/dmc_host.cpp: branch weights,
    WalkerControl::branch, computeCurData all-reduce over NCCL, swapWalkersSimple with packed device-resident walkers over
    ncclSend / ncclRecv).  The local energy of the harness is the kinetic energy (the Hamiltonian is out of scope), so the
    numbers characterise the machinery, not physics.  Returns a dict; the timed region covers `generations` generations
    with every collective and every walker transfer inside it."""
    from qmcpack_b200 import workload, sharding
    cap = int(nw * 1.5) + 8
    R = workload.initial_positions(s, cap, seed=11 + 100003 * rank)
    cr = api.Crowd(s, nw=cap, delay_rank=k, spo=spo)
    cr.set_positions(R)
    cr.mw_recompute()
    cr.vmc_init(tau=tau, use_drift=True, seed=3000 + 7919 * rank, use_cuda_graph=True, dmc=True)
    cr.set_num_walkers(nw)
    comm = api.TorchComm(dist, cr.walker_bytes, torch.device("cuda", local_rank)) if (dist is not None and world > 1) else None
    drv = api.DMCDriver(cr, tau, world * nw, branch_seed=77 + rank, comm=comm, warmup_steps=1 << 30)
    it = 0
    for _ in range(warm):
        drv.step(it)
        it += 1
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sent0 = comm.bytes_sent if comm else 0
    moves, pops, sent, t0 = 0, [], 0, time.perf_counter()
    for _ in range(generations):
        local_before = cr.nw
        ens = drv.step(it)
        it += 1
        moves += local_before * N
        pops.append(int(ens["population"]))
        sent += int(ens["walkers_sent"])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    dt_max = sharding.max_over_ranks(dt, dist, device="cuda")
    tot = sharding.reduce_block_estimator([float(moves), float(sent), float(comm.bytes_sent - sent0 if comm else 0)], dist, device="cuda")
    ke = cr.mw_block_estimators()[1]
    out = {"value": float(tot[0]) / dt_max, "unit": UNIT, "generations": generations, "ms_per_generation": 1e3 * dt_max / generations,
           "tau": tau, "target_walkers": world * nw, "population": pops, "e_trial": float(drv.history[-1]["e_trial"]),
           "walker_messages": int(tot[1]), "nccl_p2p_bytes": int(tot[2]), "walker_bytes": int(cr.walker_bytes),
           "finite": bool(np.isfinite(ke).all()),
           "path": "qmcb_vmc_sweep(dmc = 1) + qmcb_dmc_step (C++ WalkerControl::branch; all-reduce of curData and ncclSend / "
                   "ncclRecv of packed device-resident walkers through torch.distributed inside the timed region)"}
    del drv, cr
    return out


def run_b200(args, rank, local_rank, world):
    import torch
    from qmcpack_b200 import api, workload, build
    build.build()
    torch.cuda.set_device(local_rank)
    api.init(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    c = workload.CONFIGS[args.config]
    N, k, nw = c["N"], c["k"], args.walkers
    cplx = bool(c.get("complex_orbitals"))
    s = workload.make_system(N=N, M=c["M"], dtype=c["dtype"], complex_orbitals=cplx)
    lat = np.asarray(s["lattice"])
    G = np.linalg.inv(lat)
    n = N // 2
    if cplx:
        up = api.SplineSPOSet(s["coefs"][0], n, G, kind=api.C2C, kcart=s["kpts"][0])
        dn = api.SplineSPOSet(s["coefs"][1], n, G, kind=api.C2C, kcart=s["kpts"][1])
    else:
        up = api.SplineSPOSet(s["coefs"][0], n, G)
        dn = api.SplineSPOSet(s["coefs"][1], n, G)
    spo = (up, dn)
    nc = 2 if cplx else 1  # real components per orbital value
    R = workload.initial_positions(s, nw, seed=7 + 100003 * rank)  # every rank owns its own walkers (weak scaling)

    # ---------------- device-resident sweep (value)
    ndc = max(1, min(args.device_crowds, nw))
    dsizes = [nw // ndc + (1 if i < nw % ndc else 0) for i in range(ndc)]
    dcrowds, streams, off = [], [], 0
    for i in range(ndc):
        cr = api.Crowd(s, nw=dsizes[i], delay_rank=k, spo=spo)
        cr.set_positions(R[off:off + dsizes[i]])
        cr.mw_recompute()
        cr.vmc_init(tau=args.tau, use_drift=True, seed=1000 + 7919 * rank + i, use_cuda_graph=True,
                    sweep_kernel=args.sweep_kernel)  # one stream per crowd
        dcrowds.append(cr)
        streams.append(torch.cuda.ExternalStream(cr.stream, device=torch.device("cuda", local_rank)))
        off += dsizes[i]
    crowd = dcrowds[0]
    state_bytes = sum(cr.device_bytes for cr in dcrowds)

    def counts():
        a = np.concatenate([cr.vmc_counts()[0] for cr in dcrowds])
        r = np.concatenate([cr.vmc_counts()[1] for cr in dcrowds])
        return a, r

    from qmcpack_b200 import sharding

    def block_estimator(acc, rej):
        """kinetic energy of every walker of this rank, reduced over ranks: the block's only collective"""
        ke_l = np.concatenate([cr.mw_block_estimators()[1] for cr in dcrowds])
        return ke_l, sharding.reduce_block_estimator([ke_l.sum(), (ke_l * ke_l).sum(), float(nw), float(acc), float(rej)], dist,
                                                     device="cuda")
    for _ in range(max(args.warmup, 3)):
        for cr in dcrowds:
            cr.vmc_sweep_async()
    for cr in dcrowds:
        cr.sync()
    block_estimator(0, 0)  # (warm-up of the estimator kernels and of the NCCL communicator)
    a0, r0 = counts()
    launches0 = api.kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    joins = [torch.cuda.Event() for _ in range(ndc - 1)]
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local_rank) as clk:
        e0.record(streams[0])
        for st_i in streams[1:]:
            st_i.wait_event(e0)
        for _ in range(args.steps):
            for cr in dcrowds:
                cr.vmc_sweep_async()
        for ev, st_i in zip(joins, streams[1:]):
            ev.record(st_i)
            streams[0].wait_event(ev)
        # end of the block: estimator + all-reduce, inside the timed region
        a1, r1 = counts()
        ke, est = block_estimator((a1 - a0).sum(), (r1 - r0).sum())
        e1.record(streams[0])
        for cr in dcrowds:
            cr.sync()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = api.kernel_launch_count() - launches0
    acc_rate = float((a1 - a0).sum() / max(1, ((a1 - a0) + (r1 - r0)).sum()))
    lp = np.concatenate([cr.mw_block_estimators()[0] for cr in dcrowds])
    ms_max = sharding.max_over_ranks(ms, dist, device="cuda")
    value = world * nw * N * args.steps / (ms_max * 1e-3)
    ke_mean = float(est[0] / est[2])
    sane = bool(np.isfinite(ke).all() and np.isfinite(lp).all())
    sweep_kernel = dcrowds[0].sweep_kernel
    # per-kernel attribution of one sweep outside the graph (event pair around every launch)
    prof = dcrowds[0].vmc_profile_sweep()

    # ---------------- the same sweep as a caller of qmcb_vmc_sweep sees it: one call per sweep, then the per-sweep
    # results an estimator needs (kinetic energy, log psi, positions) read back to host memory every step
    dd_steps = max(2, min(args.steps, 4))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d2h_dd = 0
    for _ in range(dd_steps):
        for cr in dcrowds:
            cr.vmc_sweep_async()
        for cr in dcrowds:
            lp_i, ke_i, _, _ = cr.mw_evaluateGL()
            pos_i = cr.positions()
            d2h_dd += lp_i.nbytes + ke_i.nbytes + pos_i.nbytes
    torch.cuda.synchronize()
    dt_dd = sharding.max_over_ranks(time.perf_counter() - t0, dist, device="cuda")
    e2e_dd = {"value": world * nw * N * dd_steps / dt_dd, "unit": UNIT, "h2d_bytes_per_step": 0,
              "d2h_bytes_per_step": int(d2h_dd // dd_steps), "ms_per_step": 1e3 * dt_dd / dd_steps,
              "path": "qmcb_vmc_sweep (device-resident Metropolis loop) + per-sweep read-back of kinetic energy, log psi and "
                      "positions; informational, the contract's e2e is the host-driven path"}

    # ---------------- rank-k Woodbury flush alone (the dense contraction; informational second roofline object).
    # Executed flops per flush and walker: 4 k n^2 + 2 n k^2 (x4 complex); one-pass bytes 2 n^2 sizeof(VT).
    # Full precision runs on the FP64 tensor pipe (DMMA), mixed precision on tcgen05 TF32 with the 3-product split (3x the
    # algorithmic flops executed).
    pk = peaks()
    peak = pk["hbm_gbs"] if pk else 6650.0

    def flush_object(cr, nwf, n_, k_, cplx_, vt_bytes, label):
        us_f = cr.det_time_update_inv_mat(0, k_, 8)
        fl = (4 if cplx_ else 1) * (4.0 * k_ * n_ * n_ + 2.0 * n_ * k_ * k_) * nwf
        by = 2.0 * n_ * n_ * vt_bytes * nwf
        return {"shape": label, "walkers": nwf, "n": n_, "delay_rank": k_, "us_per_flush": us_f,
                "tflops_algorithmic": fl / us_f * 1e-6, "one_pass_GBps": by / us_f * 1e-3, "frac_hbm": by / us_f * 1e-3 / peak}
    flush = None
    try:
        vt_bytes = (4 if c["dtype"] == np.float32 else 8) * (2 if cplx else 1)
        flush = flush_object(dcrowds[0], dsizes[0], n, k, cplx, vt_bytes, args.config)
        flush["kernel"] = ("wb64::woodbury_flush_dmma_kernel (DMMA m8n8k4, one pass)" if c["dtype"] != np.float32 else
                           "wb5::woodbury_flush_tc5_kernel (tcgen05 TF32 x3 split, one pass)")
    except Exception as ex:  # measurement hook only
        flush = {"error": str(ex)}

    # ---------------- mw_recompute's FP64 inverse + log-determinant (DiracMatrixInverterCUDA::mw_invertTranspose,
    # DiracMatrixInverterCUDA.hpp:306-369) of the benchmark crowd's own Slater matrices: this library's blocked
    # Gauss-Jordan kernels (csrc/inverse.cuh) beside cublasDgetrfBatched + getriBatched, the routines the reference calls
    # (detail/CUDA/cuBLAS_LU.cu:61-210).  2 n^3 flops per matrix either way.
    recompute = None
    if rank == 0 and not args.no_recompute:
        try:
            nwr = dsizes[0]
            us_own = dcrowds[0].det_time_inverse(0, 2, reps=2)
            us_cub = dcrowds[0].det_time_inverse(0, 1, reps=1)
            fl = (4 if cplx else 1) * 2.0 * n ** 3 * nwr
            recompute = {"walkers": nwr, "n": n, "us_own": us_own, "us_cublas": us_cub, "speedup_vs_cublas": us_cub / us_own,
                         "tflops_own": fl / us_own * 1e-6, "tflops_cublas": fl / us_cub * 1e-6,
                         "kernel": "gj::gj_panel_kernel + gj::gj_update_kernel (blocked Gauss-Jordan, DMMA m8n8k4 rank-b updates)",
                         "baseline": "cublas%sgetrfBatched + getriBatched" % ("Z" if cplx else "D")}
        except Exception as ex:  # measurement hook only
            recompute = {"error": str(ex)}

    # ---------------- FP64 peak of this GPU (cublasDgemm 8192^3) and the full-precision flush at
    # the NiO-a64 (real, k = 32) and NiO-a128 (complex, k = 64) determinant shapes.  Only the determinant engine is
    # exercised: tiny tables, the benchmarked matrix sizes.
    fp64_peak, flush_fp64 = None, None
    if rank == 0 and not args.no_fp64:
        try:
            ga = torch.randn((8192, 8192), device="cuda", dtype=torch.float64)
            gb = torch.randn((8192, 8192), device="cuda", dtype=torch.float64)
            torch.matmul(ga, gb)
            best = 1e9
            for _ in range(3):
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                gc = torch.matmul(ga, gb)
                g1.record()
                g1.synchronize()
                best = min(best, g0.elapsed_time(g1))
            fp64_peak = {"tflops": 2.0 * 8192 ** 3 / (best * 1e-3) * 1e-12, "how": "torch.matmul float64 8192^3 (cublasDgemm), best of 3, CUDA events"}
            del ga, gb, gc
            flush_fp64 = []
            for label, n_, k_, cplx_, nwf in (("NiO-a64 real FP64", 384, 32, False, 512), ("NiO-a128 complex FP64", 768, 64, True, 128)):
                t = workload.random_table((4, 4, 4), n_ * (2 if cplx_ else 1), np.float64, seed=1)
                tsys = dict(n_up=n_, n_dn=n_, lattice=np.eye(3) * 4.0, coefs=[t, t])
                if cplx_:
                    kp = np.tile([0.1, 0.2, 0.3], (n_, 1))
                    tsys["kpts"] = [kp, kp]
                crf = api.Crowd(tsys, nw=nwf, delay_rank=k_)
                o = flush_object(crf, nwf, n_, k_, cplx_, 16 if cplx_ else 8, label)
                o["frac_fp64_peak"] = o["tflops_algorithmic"] / fp64_peak["tflops"]
                o["kernel"] = "wb64::woodbury_flush_dmma_kernel + binv_v_dmma_kernel (DMMA m8n8k4)"
                flush_fp64.append(o)
                del crf
        except Exception as ex:
            flush_fp64 = {"error": str(ex)}

    # ---------------- spline gather kernel alone (SPOSet::mw_evaluateVGLandDetRatioGrads), at the full population and at
    # the launch size the two-kernel sweep uses per crowd
    T = np.float32 if c["dtype"] == np.float32 else np.float64
    tdt = torch.float32 if T == np.float32 else torch.float64
    nsets = 24
    gen = torch.Generator(device="cuda").manual_seed(5 + rank)
    pos = (torch.rand((nsets, nw, 3), generator=gen, device="cuda", dtype=torch.float64) @
           torch.tensor(lat, device="cuda")).to(tdt).contiguous()
    inv = torch.randn((nw, n * nc), generator=gen, device="cuda", dtype=tdt).contiguous()
    phi = torch.empty((5, nw, n * nc), device="cuda", dtype=tdt)
    rg = torch.empty((nw, up.rg_parts, 4 * nc), device="cuda", dtype=tdt)
    ts = torch.cuda.Stream()
    lib = api.lib()

    def time_gather(nwl):
        def spline_launch(i):
            rc = lib.qmcb_spline_mw_vgl_ratio_grads_dev(up.h, nwl, pos[i % nsets].data_ptr(), inv.data_ptr(), n,
                                                        phi.data_ptr(), rg.data_ptr(), ts.cuda_stream)
            if rc:
                raise RuntimeError(lib.qmcb_last_error().decode())
        for i in range(4):
            spline_launch(i)
        ts.synchronize()
        nrep = 40
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(ts)
        for i in range(nrep):
            spline_launch(4 + i)
        s1.record(ts)
        ts.synchronize()
        return s0.elapsed_time(s1) * 1e-3 / nrep
    npad = workload.aligned_size(T, n * nc)
    esz = np.dtype(T).itemsize
    b_spl = 64 * npad * esz + 5 * n * esz * nc + n * esz * nc  # stencil + phi_vgl write + inverse-row read
    t_spl = time_gather(nw)
    tr = ncu_traffic("spline_gather_kernel") if (args.config == "NiO-a64" and nw == 512) else None
    spline_gather = {"kernel": "spline_gather_kernel (VGL + ratio/grad)", "walkers": nw, "us_per_launch": t_spl * 1e6,
                     "achieved": b_spl * nw / t_spl / 1e9, "unit": "GB/s", "frac": b_spl * nw / t_spl / 1e9 / peak,
                     "traffic": tr["bytes_per_launch"] if tr else None, "traffic_source": tr["source"] if tr else None,
                     "evals_per_s": nw / t_spl}
    nw_crowd = max(1, nw // 2)
    if nw_crowd < nw:
        t_c = time_gather(nw_crowd)
        spline_gather["per_crowd"] = {"walkers": nw_crowd, "us_per_launch": t_c * 1e6, "frac": b_spl * nw_crowd / t_c / 1e9 / peak}
    # the roofline object = the dominant kernel of the timed sweep, measured in situ (profiled sweep, event pairs)
    if sweep_kernel == 2 and prof["segment_launches"] > 0:
        moves_per_launch = N * dsizes[0] / prof["segment_launches"]
        t_k = prof["segment_us"] * 1e-6 / prof["segment_launches"]
        alg = b_spl * moves_per_launch
        trs = ncu_traffic("walker_segment_kernel") if (args.config == "NiO-a64" and dsizes[0] == 512) else None
        roofline = {"bound": "hbm", "achieved": alg / t_k / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / t_k / 1e9 / peak,
                    "traffic": trs["bytes_per_launch"] if trs else None, "traffic_source": trs["source"] if trs else None,
                    "kernel": "walker_segment_kernel (proposal + spline gather + ratio + Metropolis test + accept + next row, "
                              "%d walkers x %d moves per launch)" % (dsizes[0], round(moves_per_launch / dsizes[0])),
                    "algorithmic_bytes_per_launch": alg, "us_per_launch": t_k * 1e6,
                    "share_of_sweep": prof["segment_us"] / prof["sweep_us"],
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy bandwidth; the file carries one HBM figure)" if pk else "fallback 6.65 TB/s"}
    else:
        roofline = {"bound": "hbm", "achieved": spline_gather["achieved"], "peak": peak, "unit": "GB/s", "frac": spline_gather["frac"],
                    "traffic": spline_gather["traffic"], "traffic_source": spline_gather["traffic_source"],
                    "kernel": spline_gather["kernel"], "algorithmic_bytes_per_launch": b_spl * nw,
                    "us_per_launch": t_spl * 1e6, "share_of_sweep": prof["gather_us"] / max(prof["sweep_us"], 1e-9),
                    "frac_in_situ": (b_spl * dsizes[0] / (prof["gather_us"] * 1e-6 / max(prof["gather_launches"], 1)) / 1e9 / peak)
                    if prof["gather_launches"] else None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if pk else "fallback 6.65 TB/s"}
    roofline["sweep_attribution_us"] = {kk: prof[kk] for kk in ("sweep_us", "segment_us", "boundary_us", "gather_us", "flush_us")}

    # ---------------- end to end through the C ABI with host buffers
    e2e = None
    if not args.no_e2e:
        # one spinning host thread per crowd (the mailbox exchanges are polled): never more crowds per rank than this
        # rank's share of the box's cores (8 ranks on a 32-core box: 4 crowds each)
        ncr = max(1, min(args.crowds, nw, max(1, (os.cpu_count() or 1) // max(world, 1))))
        base, extra = divmod(nw, ncr)
        sizes = [base + (1 if i < extra else 0) for i in range(ncr)]
        crowds, off = [], 0
        del crowd, dcrowds, streams
        for i in range(ncr):
            cr = api.Crowd(s, nw=sizes[i], delay_rank=k, spo=spo)
            cr.set_positions(R[off:off + sizes[i]])
            cr.mw_recompute()
            crowds.append(cr)
            off += sizes[i]
        drv = api.HostVMC(crowds, [2000 + 17 * rank + i for i in range(ncr)], tau=args.tau, use_drift=True)
        drv.run(1)
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        drv.run(args.steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt_max = sharding.max_over_ranks(dt, dist, device="cuda")
        h2d, d2h = drv.bytes_per_sweep()
        e2e = {"value": world * nw * N * args.steps / dt_max, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "crowds": ncr,
               "ms_per_step": 1e3 * dt_max / args.steps,
               # how the per-electron calls were served: 2 = mailbox exchanges with a resident walker-segment kernel,
               # 1 = kernel launches per call (qmcb_crowd_host_kernel)
               "host_kernel_mode": sorted({c_.host_kernel for c_ in crowds}),
               "path": "qmcb_host_vmc_run: per-electron qmcb_twf_mw_eval_grad / qmcb_ps_mw_make_move / "
                       "qmcb_twf_mw_calc_ratio_grad / qmcb_twf_mw_accept_reject with host buffers, accept test on the host"}
        e2e["device_driver"] = e2e_dd
        del drv, crowds

    # ---------------- batched DMC generations (BASELINE config 5 machinery at this run's shape): population control with the
    # all-reduce and the walker exchange inside the timed region
    dmc_obj = None
    if args.dmc or not args.no_dmc:
        try:
            dmc_obj = run_dmc_generations(api, torch, dist, s, spo, k, N, nw if args.dmc else min(nw, 128), rank, local_rank, world,
                                          args.dmc_tau, generations=max(3, args.steps if args.dmc else 3), warm=2)
        except Exception as ex:
            dmc_obj = {"error": str(ex)}

    # ---------------- CPU baseline beside it (rank 0, single-GPU run only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        res = cpu_reference_run(args.config, args, steps=12, warmup=1, nw_cpu=4 * (os.cpu_count() or 1))
        cpu = {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": res["kind"], "sample": res["sample"],
               "acceptance": res["acceptance"], "acceptance_gpu": acc_rate,
               "acceptance_agrees": bool(abs(res["acceptance"] - acc_rate) < 0.01)}

    if rank == 0:
        line = {
            "metric": METRIC.replace("NiO-a64", args.config), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if T == np.float32 else "f64", "data": "synthetic",
            "config": {"workload": workload_desc(args.config, args), "walkers_per_gpu": nw, "device_crowds": ndc, "electrons": N,
                       "delay_rank": k, "table": "random orthogonal mixtures of the lowest plane waves (workload.pw_table)",
                       "l2": "inputs larger than L2: 2 x %.0f MB spline tables + %.1f GB walker state vs 126 MB L2"
                             % (up.table_bytes / 1e6, state_bytes / 1e9),
                       "parallelism": f"walkers sharded over {world} GPU(s), no data-path collective; one all-reduce per block",
                       "acceptance": acc_rate, "ke_mean_hartree": ke_mean, "finite": sane,
                       "sweep_kernel": "walker-segment kernel (csrc/segment.cuh)" if sweep_kernel == 2 else
                       "boundary kernel + spline gather per move",
                       "timed_region": "one VMC block: %d sweeps + block estimator + all-reduce over ranks" % args.steps},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "flush": flush, "flush_fp64": flush_fp64, "recompute": recompute, "fp64_peak": fp64_peak, "spline_gather": spline_gather, "cpu_baseline": cpu,
            "dmc": dmc_obj,
        }
        if args.dmc and dmc_obj and "value" in dmc_obj:
            # --dmc: the DMC generations are the headline; the VMC sweep numbers stay in the line as `vmc`
            line["vmc"] = {"value": value, "ms_per_step": ms_max / args.steps}
            line["metric"] = args.config + " DMC electron-moves/s"
            line["value"] = dmc_obj["value"]
            line["ms_per_step"] = dmc_obj["ms_per_generation"]
            line["config"]["timed_region"] = "%d DMC generations (device move loop, local energies, branch with all-reduce and walker exchange)" % dmc_obj["generations"]
        print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
