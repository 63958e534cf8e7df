#!/usr/bin/env python
"""Headline benchmark: CDU linear-MPC closed-loop offline data generation (BASELINE.json metric
"CDU linear-MPC QP solves/sec & sim-steps/sec at 1/2/4/8 B200 vs host-CPU ref").

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU path

Workload (config 3 of BASELINE.json): the crude-distillation linear MPC of the reference
(252 states, 32 inputs, 90 outputs, horizon N = 140 -> 4480 decision variables per QP,
cdu_parameters.py:94-102) on the documented synthetic stand-in plant (CDU_Model.mat is not shipped),
driven by PRBS set-point/disturbance signals with the reference's statistics
(cdu_parameters.py:115-143).  `--traj` independent closed-loop trajectories per GPU (the reference's
OS processes, lib/linearMPC.py:786-825) advance together; one bench "step" advances every
trajectory by `--slab` simulation steps: target-selector QP -> regulator QP -> plant step
(lib/linearMPC.py:845-866), i.e. traj x slab samples of the training set per GPU per step.

Numbers on the JSON line
  value       closed-loop sim-steps/s (= regulator QP solves/s; each sim step also solves one
              target-selector QP), all ranks, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e         same metric through the reference-facing host API (ClosedLoopEngine.run with NumPy
              arrays in pinned host memory -> nnmpc_sim_run_host): host->device copies of the
              slab's set-points/disturbances and device->host copies of the generated dataset
              rows are inside the timed region
  roofline    the regulator-QP iteration GEMM (FP64 tensor cores), timed live with CUDA events
  cpu_baseline  the reference-style CPU path (oracle port: cvxopt-like dense interior point) timed
              on this host on a bounded sample
The reference arm (--impl reference) times that same CPU path with every host core.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cdu_closed_loop_sim_steps_per_s"
UNIT = "sim-steps/s"
KKT_TOL = 1e-8


# ------------------------------------------------------------------------------------------ helpers
def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thr = index, [], None, None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.thr.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def _scenarios(traj, steps, seed, gain_norm=None, r_weight=None):
    """PRBS-like set-points / disturbances with the reference's hold statistics, split into `traj`
    contiguous chunks of `steps` rows (lib/linearMPC.py:786-801) -> (traj, steps, .) arrays.
    gain_norm / r_weight: conditioning study of the synthetic stand-in (defaults: plants/cdu.py, R = 0.1 I)."""
    from industrial_nnmpc_2021_b200.plants import get_cdu_problem
    from industrial_nnmpc_2021_b200.plants import cdu as _cdu
    p = get_cdu_problem(Nsim=traj * steps, seed=seed, gain_norm=_cdu.GAIN_NORM if gain_norm is None else gain_norm)
    if r_weight is not None:
        p.R = r_weight * np.eye(p.Nu)
    sp = p.setpoints.reshape(traj, steps, p.Ny)
    ds = p.disturbances.reshape(traj, steps, p.Nd)
    return p, sp, ds


# ------------------------------------------------------------------------------------------ CPU path
def _cpu_worker(args):
    """One reference-style process: closed-loop steps of one trajectory on the CPU, every QP by
    the cvxopt-like dense interior point (oracle.qp.ipm_qp), operators re-materialised per call as
    the reference does (lib/linearMPC.py:15-20, :503)."""
    seed, nsteps, threads, horizon = args
    import threadpoolctl
    import scipy.linalg
    from oracle import linear_mpc as om, qp as oq
    from industrial_nnmpc_2021_b200 import condense
    with threadpoolctl.threadpool_limits(limits=threads):
        p, sp, ds = _scenarios(1, max(nsteps, 4000), seed)
        Aa, Ba, Qa, Ra, Ma = om.augmented_matrices_for_regulator(p.A, p.B, p.Q, p.R, p.S)
        _, Pf = om.dlqr(Aa, Ba, Qa, Ra, Ma)
        P, tq = condense.condensed_hessian(Aa, Ba, Qa, Ra, Ma, Pf, horizon)   # setup, untimed
        E = np.vstack([np.eye(p.Nu), -np.eye(p.Nu)])
        G = scipy.linalg.block_diag(*([E] * horizon))                          # dense tE (:459-460)
        ots = om.TargetSelectorOracle(A=p.A, B=p.B, C=p.C, H=p.H, Bd=p.Bd, Cd=p.Cd, usp=p.usp, Rs=p.Rs,
                                      Qs=p.Qs, ulb=p.ulb, uub=p.uub)
        x, up = p.xprior.copy(), p.uprev.copy()
        iters, t0 = [], time.time()
        for t in range(nsteps):
            ysp, d = sp[0, t][:, None], ds[0, t][:, None]
            xs, us = ots.solve(ysp, d)
            x0 = np.vstack([x - xs, up - us])
            h = np.tile(np.vstack([p.uub - us, -(p.ulb - us)]), (horizon, 1))
            useq, info = oq.ipm_qp(P, tq @ x0, G, h)
            iters.append(info["iters"])
            u = useq[:p.Nu] + us
            x = p.A @ x + p.B @ u + p.Bd @ d
            up = u
        return t0, time.time(), iters


def cpu_reference_rate(nsteps, horizon=140, seed=11):
    """sim-steps/s of the CPU path using every host core: `workers` processes (the reference's
    num_parallel) x `threads` BLAS threads each."""
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    threads = min(cores, 8)
    mem_gb = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2**30
    workers = max(1, min(cores // threads, int(mem_gb // 4), 32))
    t0 = time.perf_counter()
    if workers == 1:
        res = [_cpu_worker((seed, nsteps, threads, horizon))]
    else:
        with mp.get_context("spawn").Pool(workers) as pool:
            res = pool.map(_cpu_worker, [(seed + 2 * w, nsteps, threads, horizon) for w in range(workers)])
    wall = time.perf_counter() - t0
    loop = max(r[1] for r in res) - min(r[0] for r in res)   # first loop start -> last loop end (setup excluded)
    rate = workers * nsteps / loop
    iters = [i for r in res for i in r[2]]
    info = dict(cores=workers * threads, workers=workers, threads_per_worker=threads, host_cores=cores,
                sample=f"{workers} trajectories x {nsteps} closed-loop steps (CDU, N={horizon}, n={horizon * 32}), "
                       f"cold-start dense interior point per QP, mean {np.mean(iters):.1f} IPM iterations",
                seconds_per_qp=loop / nsteps, loop_s=loop, wall_s=wall)
    return rate, info


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r, info = cpu_reference_rate(args.ref_steps, horizon=args.horizon)
    runs = [r]
    # keep the whole arm within a few minutes whatever the host: as many of the W+K runs as fit in 240 s
    total = min(args.warmup + args.steps, max(1, int(240 // max(info["wall_s"], 1e-3))))
    for _ in range(total - 1):
        r, info = cpu_reference_rate(args.ref_steps, horizon=args.horizon)
        runs.append(r)
    rates = runs[min(args.warmup, total - 1):]
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(rates), "warmup": args.warmup, "ms_per_step": 1e3 * args.ref_steps * info["workers"] / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(args, info["workers"], args.ref_steps),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": "port",
                         "sample": info["sample"], "seconds_per_qp": info["seconds_per_qp"],
                         "host_cores": info["host_cores"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference CPU path = oracle port of lib/linearMPC.py simulate_offline with a cvxopt-like dense "
                "interior point (cvxopt itself is not installable here); paper: 35 s/QP, 3.57 steps/s on 149 procs",
    }
    print(json.dumps(line), flush=True)


def _config(args, traj, slab):
    return {"workload": "CDU linear-MPC closed-loop offline data generation (BASELINE.json configs[2]): synthetic "
                        "CDU stand-in 252x32x90, reference tuning, PRBS scenarios",
            "Nx": 252, "Nu": 32, "Ny": 90, "horizon": args.horizon, "qp_vars": args.horizon * 32,
            "trajectories_per_gpu": traj, "concurrent_slots_per_gpu": min(traj, getattr(args, "slots", traj)),
            "sim_steps_per_step": slab, "tol_kkt": KKT_TOL,
            "scenario_slabs": f"{getattr(args, 'unique_slabs', 0)} distinct PRBS slabs at most, cycled over the steps",
            "precision": args.precision,
            "l2": "no flush: operators (2 x 161 MB FP64 + 2 x 40 MB fp16) + solver state exceed the 126 MB L2 every iteration",
            "parallelism": f"trajectories sharded over {args.gpus} GPU(s), no collective on the solve path"}


# ------------------------------------------------------------------------------------------ GPU path
def run_native(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this framework has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    import __graft_entry__ as entry
    if rank == 0:
        from industrial_nnmpc_2021_b200 import build
        build.build()
    if world > 1:
        dist.barrier()
    from industrial_nnmpc_2021_b200 import _lib
    from industrial_nnmpc_2021_b200.linearMPC import LinearMPCController, ClosedLoopEngine
    L = _lib.lib()

    K, W, B, Ts = args.steps, max(args.warmup, 3), args.traj, args.slab
    nslab = 2 * (W + K)                      # device-resident phase, then the end-to-end phase
    # distinct PRBS scenario slabs kept in memory; further steps cycle through them (bounds host memory: one slab
    # is traj x slab x 95 doubles).  Trajectories that continue across steps (--traj <= --slots) see one extra
    # set-point jump where the cycle wraps.
    nuniq = min(nslab, max(1, args.unique_slabs))
    p, sp, ds = _scenarios(B, nuniq * Ts, seed=101 + 2 * rank, gain_norm=args.gain_norm, r_weight=args.r_weight)
    if args.horizon != p.N:
        p.N = args.horizon
    t_setup = time.perf_counter()
    ts = LinearMPCController.setup_target_selector(p.A, p.B, p.C, p.H, p.Bd, p.Cd, p.usp, p.Qs, p.Rs, p.ulb, p.uub,
                                                   device=dev)
    solver_kw = {}
    if args.alpha is not None:
        solver_kw["alpha"] = args.alpha
    if args.rho_scale is not None:
        solver_kw["rho_scale"] = args.rho_scale
    reg = LinearMPCController.setup_regulator(p.A, p.B, p.Q, p.R, p.S, p.N, p.ulb, p.uub, device=dev, **solver_kw)
    eng = ClosedLoopEngine(reg, ts, p.A, p.B, p.Bd, precision=args.precision, slots=args.slots)
    t_setup = time.perf_counter() - t_setup
    n, nx, nu, ny, nd = p.N * p.Nu, p.Nx, p.Nu, p.Ny, p.Nd

    f64 = dict(dtype=torch.float64, device=dev)
    slabs = [(torch.tensor(np.ascontiguousarray(sp[:, i * Ts:(i + 1) * Ts]), **f64),
              torch.tensor(np.ascontiguousarray(ds[:, i * Ts:(i + 1) * Ts]), **f64)) for i in range(min(W + K, nuniq))]
    out_d = dict(x=torch.empty((B, Ts, nx), **f64), uprev=torch.empty((B, Ts, nu), **f64),
                 xs=torch.empty((B, Ts, nx), **f64), us=torch.empty((B, Ts, nu), **f64),
                 u=torch.empty((B, Ts, nu), **f64), iters=torch.empty((B, Ts), dtype=torch.int32, device=dev),
                 kkt=torch.empty((B, Ts), **f64))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxrank(v):
        if world == 1:
            return v
        t = torch.tensor([v], **f64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident phase -------------------------------------------------------------
    x, up = p.xprior, p.uprev
    hit = False
    kkt_acc = torch.zeros((), **f64)
    it_sum = torch.zeros((), dtype=torch.int64, device=dev)
    it_acc = torch.zeros((), dtype=torch.int32, device=dev)
    for i in range(W):
        r = eng.run(x, up, *slabs[i % len(slabs)], resume=i > 0, out=out_d, max_iter=args.max_iter)
        x, up = r["x_final"], r["uprev_final"]
    barrier()
    _lib.prof_enable(True)
    _lib.prof_read(reset=True)
    st0 = eng.stats()
    launches0 = L.nnmpc_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        ev0.record()
        for i in range(W, W + K):
            r = eng.run(x, up, *slabs[i % len(slabs)], resume=True, out=out_d, max_iter=args.max_iter)
            x, up = r["x_final"], r["uprev_final"]
            hit |= bool(r["maxiter_hit"])
            # validity of the timed work itself (device-side reductions, read after the timing)
            kkt_acc = torch.maximum(kkt_acc, r["kkt"].max())
            it_sum = it_sum + r["iters"].sum(dtype=torch.int64)
            it_acc = torch.maximum(it_acc, r["iters"].max())
        ev1.record()
        barrier()
    ms_dev = maxrank(ev0.elapsed_time(ev1))
    launches = L.nnmpc_launch_count() - launches0
    chans = _lib.prof_readn(4, reset=True)
    (gemm_ms, gemm_flops, gemm_launches), (f64_ms, f64_flops, f64_launches) = chans[0], chans[1]
    tail_ms, rest_ms = chans[2][0], chans[3][0]
    _lib.prof_enable(False)
    st1 = eng.stats()
    kkt_max, it_sum, it_max = float(kkt_acc), int(it_sum), int(it_acc)
    clocks = clk.summary()
    value = world * B * Ts * K / (ms_dev * 1e-3)

    # ---- end-to-end phase: host buffers through the reference-facing API -----------------
    def pinned(shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, pin_memory=True).numpy()
    out_h = dict(x=pinned((B, Ts, nx)), uprev=pinned((B, Ts, nu)), xs=pinned((B, Ts, nx)), us=pinned((B, Ts, nu)),
                 u=pinned((B, Ts, nu)), iters=pinned((B, Ts), torch.int32), kkt=pinned((B, Ts)))
    sp_h, ds_h = pinned((B, Ts, ny)), pinned((B, Ts, nd))
    xh, uph = x.cpu().numpy(), up.cpu().numpy()
    e2e_kkt = 0.0
    barrier()
    t_e2e = None
    for i in range(W + K, 2 * (W + K)):
        if i == 2 * W + K:
            barrier()
            ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ee0.record()
            t_e2e = time.perf_counter()
        j = i % nuniq
        sp_h[...] = sp[:, j * Ts:(j + 1) * Ts]          # this step's inputs staged in pinned host memory
        ds_h[...] = ds[:, j * Ts:(j + 1) * Ts]
        r = eng.run(xh, uph, sp_h, ds_h, resume=True, out=out_h, max_iter=args.max_iter)
        xh, uph = r["x_final"], r["uprev_final"]
        e2e_kkt = max(e2e_kkt, float(out_h["kkt"].max()))      # the device->host result is read every step
        hit |= bool(r["maxiter_hit"])
    ee1.record()
    barrier()
    wall_e2e = time.perf_counter() - t_e2e
    ms_e2e = maxrank(max(ee0.elapsed_time(ee1), 1e3 * wall_e2e))
    e2e_value = world * B * Ts * K / (ms_e2e * 1e-3)
    h2d = (sp_h.nbytes + ds_h.nbytes + xh.nbytes + uph.nbytes) * world
    d2h = (sum(v.nbytes for v in out_h.values()) + xh.nbytes + uph.nbytes) * world

    # ---- dataset gather over NCCL/NVLink (outside `value`; the solve path has no collective) ----
    gather = None
    if world > 1:
        from industrial_nnmpc_2021_b200 import distributed as nd_
        local = {k: out_d[k] for k in nd_.DATASET_KEYS}
        nd_.gather_chunks(local, world * B)          # warm-up (NCCL communicator set-up)
        barrier()
        gv0, gv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gv0.record()
        full = nd_.gather_chunks(local, world * B)
        gv1.record()
        barrier()
        g_ms = maxrank(gv0.elapsed_time(gv1))
        g_bytes = sum(v.numel() * v.element_size() for v in full.values())
        gather = {"ms": g_ms, "bytes_gathered_per_rank": g_bytes, "algbw_GBs": g_bytes / (g_ms * 1e-3) / 1e9,
                  "what": "all_gather_into_tensor of one step's dataset rows (x, uprev, xs, us, u) from every rank"}
        del full

    if kkt_max > KKT_TOL or e2e_kkt > KKT_TOL or hit:
        raise SystemExit(f"bench.py: timed solves missed the tolerance (kkt {kkt_max:.2e}/{e2e_kkt:.2e}, "
                         f"maxiter_hit={hit}); number rejected")

    # ---- roofline of the dominant kernel ------------------------------------------------------
    peaks, peak_src = _peaks()
    a = torch.randn((6144, 6144), **f64); b = torch.randn((6144, 6144), **f64)
    for _ in range(3):
        torch.matmul(a, b)
    best = 0.0
    for _ in range(5):
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(); torch.matmul(a, b); g1.record(); torch.cuda.synchronize()
        best = max(best, 2 * 6144 ** 3 / (g0.elapsed_time(g1) * 1e-3) / 1e12)
    del a, b
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    step_ms_local = ms_dev if world == 1 else ev0.elapsed_time(ev1)
    dgemm_src = ("cuBLAS DGEMM 6144^3 measured in this run (FP64 has no tcgen05 kind; "
                 f"MEASURED_PEAKS.json carries bf16 {peaks.get('bf16_tflops')} TF/s and HBM "
                 f"{peaks.get('hbm_gbs')} GB/s only, {peak_src})")
    nvar = args.horizon * 32
    per_unit = f"active samples x 2 n^2 (n = {nvar}: {2 * nvar * nvar / 1e6:.2f} MFLOP per sample-iteration)"
    if args.precision == "f64":
        roofline = {"bound": "tensor", "kernel": "gemm_f64_kernel<EpiAdmm> (regulator-QP iteration, FP64 DMMA)",
                    "achieved": achieved, "peak": best, "unit": "TFLOP/s", "frac": achieved / best if best else None,
                    "traffic": None, "launches": gemm_launches, "avg_launch_ms": gemm_ms / max(gemm_launches, 1),
                    "share_of_step": gemm_ms / step_ms_local, "peak_source": dgemm_src, "flops_per_launch": per_unit}
        roofline2 = None
    else:
        lp_peak = float(peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1590.0)
        f64_ach = f64_flops / (f64_ms * 1e-3) / 1e12 if f64_ms > 0 else 0.0
        # DRAM traffic per launch, scaled from the one `ncu --set full` capture of this kernel (profiles/
        # r01af_ncu_full_lp_gemm.txt: dram read 1.2787 GB + write 0.7549 GB at 8192 rows, n = 4480): 80.3 MB of fp16
        # operator per launch + 238.4 kB per row (algorithmic: 42 B x n = 188.2 kB per row).  Only for n = 4480.
        rows_per_launch = (gemm_flops / (2.0 * nvar * nvar)) / max(gemm_launches, 1)
        lp_traffic = (80.3e6 + 238.4e3 * rows_per_launch) if nvar == 4480 else None
        r_lp = {"bound": "tensor", "kernel": "lp_gemm_kernel<EpiDelta> (regulator-QP iteration: tcgen05 kind::f16, fp16 "
                                             "increments x two-term fp16 operator split, fp32 TMEM accumulators, FP64 state)",
                "achieved": achieved, "executed_mma": 2.0 * achieved, "peak": lp_peak, "unit": "TFLOP/s",
                "frac": achieved / lp_peak, "frac_executed": 2.0 * achieved / lp_peak, "traffic": lp_traffic,
                "traffic_algorithmic": 80.3e6 + 188.2e3 * rows_per_launch if nvar == 4480 else None,
                "rows_per_launch": rows_per_launch,
                "launches": gemm_launches, "avg_launch_ms": gemm_ms / max(gemm_launches, 1),
                "share_of_step": gemm_ms / step_ms_local,
                "peak_source": f"dense 16-bit tensor throughput, sustained figure of MEASURED_PEAKS.json ({peak_src}); "
                               "kernel timed inside a long step",
                "flops_per_launch": per_unit + "; the kernel executes 2x that in MMA work (T1 and T2 products)"}
        if os.environ.get("NNMPC_EXACT_GEMM", "int8") == "dmma":
            r_64 = {"bound": "tensor", "kernel": "gemm_f64_kernel<EpiAnchor|EpiVerifyMax> (FP64 anchors x = Top w - c and "
                                                 "exact KKT checks P z + q, DMMA, small row lists)",
                    "achieved": f64_ach, "peak": best, "unit": "TFLOP/s", "frac": f64_ach / best if best else None,
                    "traffic": None, "launches": f64_launches, "avg_launch_ms": f64_ms / max(f64_launches, 1),
                    "share_of_step": f64_ms / step_ms_local, "peak_source": dgemm_src,
                    "flops_per_launch": "listed samples x 2 n^2"}
        else:
            # INT8-sliced exact applies: 21 (anchors, 6 levels) / 36 (checks, 8 levels) INT8 products of 2 n^2 ops per row
            n_anch, n_chk = st1["anchors"] - st0["anchors"], st1["exact_checks"] - st0["exact_checks"]
            int8_ops = 2.0 * nvar * nvar * (21.0 * n_anch + 36.0 * n_chk)
            int8_ach = int8_ops / (f64_ms * 1e-3) / 1e12 if f64_ms > 0 else 0.0
            int8_peak = 2.0 * lp_peak
            r_64 = {"bound": "tensor", "kernel": "oz_gemm2_kernel<.,.,128,OzEpiAnchor|OzEpiVerify> + k_oz_slice (FP64-accurate "
                                                 "anchors x = Top w - c and exact KKT checks P z + q on tcgen05 kind::i8: "
                                                 "error-free base-128 digit planes, exact INT32 accumulation)",
                    "achieved": int8_ach, "fp64_equivalent": f64_ach, "peak": int8_peak, "unit": "TOP/s",
                    "frac": int8_ach / int8_peak, "fp64_equivalent_vs_cublas_dgemm": f64_ach / best if best else None,
                    "cublas_dgemm_tflops": best,
                    "traffic": None, "launches": f64_launches, "avg_launch_ms": f64_ms / max(f64_launches, 1),
                    "share_of_step": f64_ms / step_ms_local,
                    "peak_source": "2 x the sustained 16-bit dense figure of MEASURED_PEAKS.json (INT8 dense is nominally twice "
                                   "the 16-bit rate; the file carries no measured INT8 entry)",
                    "flops_per_launch": "listed samples x 2 n^2 FP64-equivalent = x 21 (anchor) or 36 (check) INT8 products; "
                                        "span includes the digit-plane slicing kernel"}
        roofline, roofline2 = (r_lp, r_64) if gemm_ms >= f64_ms else (r_64, r_lp)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, info = cpu_reference_rate(1, horizon=args.horizon)
            cpu = {"value": rate, "unit": UNIT, "cores": info["cores"], "kind": "port", "sample": info["sample"],
                   "seconds_per_qp": info["seconds_per_qp"], "host_cores": info["host_cores"]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": _config(args, B, Ts),
            "qp_solves_per_s": {"regulator": value, "target_selector": value},
            "iterations": {"mean": it_sum / (B * Ts * K), "max": it_max, "kkt_max": kkt_max},
            "conditioning": {"cond_P": reg.eig_range[1] / reg.eig_range[0], "lambda_min": reg.eig_range[0],
                             "lambda_max": reg.eig_range[1],
                             "active_bound_frac": (st1["qps_with_active_bounds"] - st0["qps_with_active_bounds"])
                             / max(st1["qps"] - st0["qps"], 1),
                             "mean_active_bounds_per_qp": (st1["active_bounds"] - st0["active_bounds"])
                             / max(st1["qps"] - st0["qps"], 1),
                             "gain_norm": args.gain_norm, "R": args.r_weight,
                             "what": "condition number of the condensed Hessian; share of the timed QPs whose optimum "
                                     "has >= 1 active input bound; active bounds per QP (of n)"},
            "roofline": roofline, "roofline_second_kernel": roofline2, "cpu_baseline": cpu,
            "precision": args.precision,
            "solver_work_per_qp": {k: (st1[k] - st0[k]) / max(st1["qps"] - st0["qps"], 1)
                                   for k in ("row_iterations", "anchors", "exact_checks")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / K, "kkt_max": e2e_kkt},
            "time_breakdown": {"unit": "share of the timed region (CUDA events on the launching stream, rank 0)",
                               "iteration_passes": gemm_ms / step_ms_local, "exact_anchors_and_checks": f64_ms / step_ms_local,
                               "fp64_tail_iterations": tail_ms / step_ms_local,
                               "plant_target_qbuild_lists": rest_ms / step_ms_local,
                               "other": 1.0 - (gemm_ms + f64_ms + tail_ms + rest_ms) / step_ms_local},
            "gpu_launches": int(launches), "clocks": clocks, "setup_s": t_setup, "gather": gather,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--unique-slabs", type=int, default=6,
                    help="distinct scenario slabs generated; steps beyond that cycle through them")
    ap.add_argument("--traj", type=int, default=65536,
                    help="closed-loop trajectory chunks per GPU and bench step (the reference's per-process chunks, "
                         "lib/linearMPC.py:786-801); with --slots below it they queue up (continuous batching)")
    ap.add_argument("--slab", type=int, default=16, help="simulation steps every trajectory advances per bench step")
    ap.add_argument("--horizon", type=int, default=140)
    ap.add_argument("--ref-steps", type=int, default=1, help="closed-loop steps per worker per reference step")
    ap.add_argument("--slots", type=int, default=16384,
                    help="trajectories advanced concurrently per GPU; with --traj above it the other chunks queue up and "
                         "finished slots take the next one (continuous batching)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="mixed", choices=["mixed", "f64"],
                    help="regulator-QP iteration arithmetic: tcgen05 fp16 increments + FP64 anchors, or all FP64 DMMA")
    ap.add_argument("--alpha", type=float, default=None, help="Douglas-Rachford relaxation (solver default 1.8)")
    ap.add_argument("--rho-scale", type=float, default=None, help="multiplier of the default ADMM penalty (solver default 1)")
    ap.add_argument("--gain-norm", type=float, default=None,
                    help="conditioning study: steady-state gain-row norm of the synthetic plant (default plants/cdu.py: 0.7)")
    ap.add_argument("--r-weight", type=float, default=None, help="conditioning study: R = r I (reference tuning: 0.1)")
    ap.add_argument("--max-iter", type=int, default=3000, help="per-QP iteration cap (a hit rejects the number)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
