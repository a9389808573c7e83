#!/usr/bin/env python
"""bench.py -- QP-subproblem hot path of SqpSolver.jl on B200.

Workload (BASELINE.json configs[4], the configuration the 1/2/4/8-GPU metric is quoted on):
a batch of 1024 perturbed-load copies of the case118-shaped synthetic ACOPF network
(pd, qd x (1 + 0.05 N(0,1)), Philox key 1234 / stream = instance id), one shared sparsity
pattern, sharded in contiguous blocks over the ranks (strong scaling: the batch is fixed).

A "step" is one SQP iteration of the hot path over the rank's shard: COO value scatter (K2),
the batched QP-subproblem solve (ADMM + PCG + polish kernel), the merit / violation norms and
the KT residual.  The per-step inputs are the *real* ones: an untimed set-up phase runs the
batched SQP-TR driver on the device and records, for its first rounds, what the host handed
to the engine (dE, h_val, df, E, x_k, Delta, active mask); the timed steps replay those rounds
in order, including the warm start carried from round to round.

  value : device-resident inputs (update_nlp_device + solve_tr_device), CUDA events on the
          engine's stream, L2 flushed between steps, max over ranks.
  e2e   : the same steps through the host C-ABI calls a SqpSolver.jl user makes
          (update_nlp / solve_tr / merit / kt_residuals with HOST buffers: pinned staging,
          H2D and D2H inside the timed region).
  --impl reference : the CPU restatement of the reference path (oracle SQP-TR + interior
          point QP solver standing in for Ipopt) on the host cores, same metric and workload.
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
sys.path.insert(0, ROOT)

METRIC = "sqp_iterations_per_sec"
UNIT = "SQP iterations/s (= QP subproblem solves/s, summed over the instances of the batch)"


def make_workload(args):
    from sqpsolver_jl_b200.nlp.networks import case9, synth_net

    if args.workload == "batch118":
        return synth_net(118, 186, 54, seed=118), args.batch, "batch of %d perturbed-load case118-shaped ACOPF instances" % args.batch
    if args.workload == "batch9":
        return case9(), args.batch, "batch of %d perturbed-load case9 ACOPF instances" % args.batch
    raise SystemExit("unknown workload " + args.workload)


def sqp_params(args):
    return dict(max_iter=args.sqp_max_iter, init_mu=args.init_mu, use_soc=False)


# ------------------------------------------------------------------ bytes model
def bytes_model(n, m, nnzJ, nnzH):
    spmv = lambda rows, cols, nnz: 12 * nnz + 4 * (rows + 1) + 8 * cols + 8 * rows
    b_cg = spmv(n, n, nnzH) + spmv(m, n, nnzJ) + spmv(n, m, nnzJ) + 8 * (10 * n + 2 * m)
    b_admm = spmv(m, n, nnzJ) + spmv(n, m, nnzJ) + 8 * (8 * n + 10 * m)
    b_check = spmv(m, n, nnzJ) + spmv(n, n, nnzH) + 2 * spmv(n, m, nnzJ) + 8 * (6 * n + 6 * m)
    return b_cg, b_admm, b_check


def bytes_model_ipm(n, m, nnzJ, nnzH, nnzL):
    """Algorithmic bytes of the interior-point path: per-instance VALUE bytes each phase must
    touch once (fp64); the int32 index programs are shared by the batch and counted once per launch."""
    resid = 8 * (nnzH + nnzJ) + 8 * (4 * n + 3 * m)            # P x, J' lam, r_x, side residual norms
    assemble = 8 * (nnzJ + nnzH + m + n) + 8 * nnzL            # K = P + D + J'WJ  -> L
    factor = 16 * nnzL                                         # read K, write L
    solve = 16 * nnzL + 24 * n                                 # forward + backward sweep
    rhs_step = 8 * (2 * nnzJ) + 8 * (8 * n + 14 * m)           # J't, J dx, step ratio, update of x,s,z,y,r
    return resid + rhs_step + solve, assemble + factor         # (per iteration, per factorisation)


def index_bytes_ipm(chol, nnzJ, nnzH, n, m):
    return 4 * (2 * chol["flops"] // 2 + 6 * chol["nnzL"] + 2 * (nnzJ + nnzH) + 4 * (n + m))


def algorithmic_bytes(info, sel, n, m, nnzJ, nnzH, chol=None):
    b_cg, b_admm, b_check = bytes_model(n, m, nnzJ, nnzH)
    i = info[sel]
    tot = float(((i["cg_iters"].astype(np.int64) + i["polish_cg_iters"]) * b_cg + i["admm_iters"].astype(np.int64) * b_admm
                 + i["checks"].astype(np.int64) * b_check).sum())
    if chol and chol["nnzL"]:
        b_it, b_f = bytes_model_ipm(n, m, nnzJ, nnzH, chol["nnzL"])
        tot += float((i["ipm_iters"].astype(np.int64) * b_it + i["chol_factorizations"].astype(np.int64) * b_f).sum())
        tot += index_bytes_ipm(chol, nnzJ, nnzH, n, m)
    return tot


def spmv_roofline(local, dev, peak, batch, reps=10):
    """SpMV leg of BASELINE.json's metric: the three CSR products of the path (J p, J' lambda, H p;
    sqp_trust_region.jl:343,490,492, common.jl:17) through csrc/spmv.cuh on the ~2000-bus synthetic network
    (BASELINE configs[3]), batched so that the value streams (batch x nnz x 8 B) exceed the 126 MB L2.
    achieved = algorithmic bytes (fp64 values + x + y per instance, int32 pattern once) / CUDA-event time."""
    import torch
    from sqpsolver_jl_b200 import capi
    from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
    from sqpsolver_jl_b200.nlp.networks import synth_net

    net = synth_net(2000, 3000, 400, seed=2000)
    nlp = AcopfPolar(net)
    eng = capi.Engine(local)
    try:
        eng.setup_nlp(nlp.n, nlp.m, nlp.num_linear_constraints, nlp.j_row, nlp.j_col, nlp.h_row, nlp.h_col,
                      nlp.x_L, nlp.x_U, nlp.g_L, nlp.g_U, batch=batch)
        rng = np.random.default_rng(2000)
        dE = rng.standard_normal((1, nlp.nnz_jac_coo)) + 1e-3 * rng.standard_normal((batch, 1))
        hv = rng.standard_normal((1, nlp.nnz_hess_coo)) + 1e-3 * rng.standard_normal((batch, 1))
        eng.update_nlp(dE, hv, np.zeros((batch, nlp.n)), np.zeros((batch, nlp.m)))
        nnz = {0: int(eng.get_csr(0)[1].shape[0]), 2: int(eng.get_csr(2)[1].shape[0])}
        nnz[1] = nnz[0]
        dims = {0: (nlp.m, nlp.n), 1: (nlp.n, nlp.m), 2: (nlp.n, nlp.n)}
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        out = {}
        for which, name in ((0, "J*x"), (1, "J'*y"), (2, "H*x")):
            rows, cols = dims[which]
            x = torch.randn(batch, cols, dtype=torch.float64, device=dev)
            y = torch.empty(batch, rows, dtype=torch.float64, device=dev)
            ms = []
            for it in range(3 + reps):
                flush.fill_(it & 0xFF)
                torch.cuda.synchronize()
                eng.spmv_device(which, x.data_ptr(), y.data_ptr())
                t = eng.last_solve_ms  # events recorded on the engine's stream around the launch
                if it >= 3:
                    ms.append(t)
            byt = batch * 8 * (nnz[which] + rows + cols) + 4 * nnz[which] + 8 * rows
            t_avg = float(np.mean(ms))
            out[name] = {"rows": rows, "cols": cols, "nnz": nnz[which], "ms": t_avg, "GBps": byt / (t_avg * 1e-3) / 1e9,
                         "frac": byt / (t_avg * 1e-3) / 1e9 / peak if peak else None}
        return {"workload": "ACOPF ~2000-bus synthetic network (n=%d, m=%d), %d instances with one shared pattern" % (nlp.n, nlp.m, batch),
                "kernel": "k_spmv_stream (CSR-stream, csrc/spmv.cuh)", "reps": reps, "l2": "256 MiB flush before every launch",
                "peak_GBps": peak, "products": out}
    finally:
        eng.close()


def single_instance_2000(local, peak, max_iter=8):
    """BASELINE configs[3]: the ~2000-bus synthetic network as ONE instance on one GPU (cooperative-grid team, k_solve_grid):
    the QP / restoration subproblems of the first SQP iterations of the trust-region run -- the solve the reference issues
    at subproblem_JuMP.jl:178 -- timed with CUDA events around every solve launch.  achieved = algorithmic bytes (the
    interior-point bytes model of DESIGN.md section 5 with this network's sizes x the device-counted iterations and
    factorisations) / kernel time.  The working set (~10 MB of values + the index programs) is L2-resident, so this is a
    latency figure stated against the HBM roofline, as SURVEY section 8d asks; traffic = dram bytes of one launch from
    the ncu capture recorded in profiles/r02_traffic_2000.json when it matches this kernel."""
    from sqpsolver_jl_b200 import capi
    from sqpsolver_jl_b200.host.sqp_trust_region import Parameters, SqpTR
    from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
    from sqpsolver_jl_b200.nlp.networks import synth_net

    nlp = AcopfPolar(synth_net(2000, 3000, 400, 2000))
    t0 = time.perf_counter()
    d = SqpTR(nlp, Parameters(max_iter=max_iter, init_mu=1e5), device=local)
    setup_s = time.perf_counter() - t0
    eng = d.batch.optimizer.engine
    per = []
    orig = d.batch.optimizer._solve

    def hook(phase, x_k, delta, E_override=None, active=None):
        r = orig(phase, x_k, delta, E_override, active)
        info = d.batch.optimizer.last_info[0]
        per.append((int(phase), float(eng.last_solve_ms), int(info["ipm_iters"]), int(info["chol_factorizations"]), int(info["moi_status"])))
        return r

    d.batch.optimizer._solve = hook
    t0 = time.perf_counter()
    d.run()
    wall = time.perf_counter() - t0
    chol = eng.chol_stats()
    nnzJ, nnzH = int(eng.get_csr(0)[1].shape[0]), int(eng.get_csr(2)[1].shape[0])
    kernel = eng.last_solve_kernel
    d.close()
    sub = [q for q in per if q[0] in (capi.PHASE_QP, capi.PHASE_FR)]
    b_it, b_f = bytes_model_ipm(nlp.n, nlp.m, nnzJ, nnzH, max(chol["nnzL"], 1))
    ms = float(sum(q[1] for q in sub))
    its = int(sum(q[2] for q in sub))
    nf = int(sum(q[3] for q in sub))
    alg = its * b_it + nf * b_f
    ach = alg / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic_2000.json")))
        if tr.get("kernel") == kernel:
            traffic = tr
    except (OSError, ValueError):
        pass
    return {"workload": "ACOPF ~2000-bus synthetic network, one instance (n=%d, m=%d, nnzJ=%d, nnzH_sym=%d)" % (nlp.n, nlp.m, nnzJ, nnzH),
            "kernel": kernel, "chol": chol, "subproblems": len(sub), "ms_per_qp": ms / max(1, len(sub)),
            "qp_solves_per_sec": len(sub) / (ms * 1e-3) if ms > 0 else None, "ms_per_ipm_iteration": ms / max(1, its),
            "ipm_iterations": its, "factorizations": nf, "sqp_iterations_per_sec_wall": len(sub) / wall if wall > 0 else None,
            "setup_s_symbolic_analysis": setup_s,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak else None,
                         "bytes_per_ipm_iteration": b_it, "bytes_per_factorization": b_f,
                         "traffic": traffic["dram_bytes"] if traffic else None,
                         "traffic_source": {k: traffic[k] for k in ("profile", "launch", "git") if k in traffic} if traffic else None,
                         "note": "one instance: the working set is L2-resident and the kernel is bound by grid-barrier and dependent-"
                                 "load latency (DESIGN.md 5.3), the fraction is reported because BASELINE asks for it"},
            "phases": [{"phase": q[0], "ms": q[1], "ipm_iters": q[2], "status": q[4]} for q in per]}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self, wait_s=3.0):
        """Starts the sampler and waits for its first row: the start-up of nvidia-smi (NVML initialisation) must not fall into the
        timed region -- it can hold up the thread that enqueues the launches, and the GPU runs dry."""
        self._start()
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < wait_s:
            time.sleep(0.02)

    def _start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ reference arm / cpu baseline
def _oracle_qp_worker(job):
    """Solve one recorded QP subproblem with the CPU oracle; returns seconds."""
    from oracle import qp_solver as qs
    from oracle.coo import CooMatrix, SymCooMatrix
    from oracle.subproblem import trust_region_box
    from sqpsolver_jl_b200.nlp.acopf import AcopfPolar

    net, pd, qd, rec = job
    nlp = AcopfPolar(net, pd=pd, qd=qd)
    J = CooMatrix(nlp.j_row, nlp.j_col, nlp.m, nlp.n); J.fill(rec["dE"])
    H = SymCooMatrix(nlp.h_row, nlp.h_col, nlp.n); H.fill(rec["h_val"])
    lb, ub = trust_region_box(nlp.x_L - rec["x"], nlp.x_U - rec["x"], rec["Delta"])
    t0 = time.perf_counter()
    res = qs.solve_qp(H.to_scipy(), rec["df"], J.to_scipy(), nlp.g_L - rec["E"], nlp.g_U - rec["E"], lb, ub)
    return time.perf_counter() - t0, res.status


def _oracle_sqp_worker(job):
    """Run `iters` SQP iterations of one instance with the CPU oracle; returns (#QP solves, seconds)."""
    from oracle.sqp_tr import Parameters, SqpTROracle
    from sqpsolver_jl_b200.nlp.acopf import AcopfPolar

    net, pd, qd, kw, iters = job
    kw = dict(kw, max_iter=iters)
    t0 = time.perf_counter()
    s = SqpTROracle(AcopfPolar(net, pd=pd, qd=qd), Parameters(**kw)).run()
    return int(s.iter) - 1 if s.ret == -1 else int(s.iter), s.n_qp, time.perf_counter() - t0


def run_reference(args):
    """`--impl reference`: the CPU path on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    net, batch, wname = make_workload(args)
    cores = os.cpu_count() or 1
    ninst = min(batch, 2 * cores)   # bounded sample of the batch: two instances per core and step
    pd, qd = net.perturbed_loads(ninst)
    kw = sqp_params(args)
    # one "step" = one SQP iteration on each of `ninst` instances, `cores` at a time (bounded sample of the batch)
    with mp.Pool(cores) as pool:
        jobs = lambda iters: [(net, pd[b], qd[b], kw, iters) for b in range(ninst)]
        if args.warmup:
            pool.map(_oracle_sqp_worker, jobs(min(args.warmup, 2)))
        t0 = time.perf_counter()
        out = pool.map(_oracle_sqp_worker, jobs(args.steps))
        wall = time.perf_counter() - t0
    iters = sum(o[0] for o in out)
    value = iters / wall
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wname, "network": {"nbus": net.nbus, "nbranch": net.nbranch, "ngen": net.ngen, "seed": net.meta.get("seed")},
                   "batch_total": batch, "sqp": kw,
                   "step": "one SQP iteration (evaluation + QP subproblem solve) on each sampled instance",
                   "sample": f"{ninst} of the {batch} instances x {args.steps} SQP iterations", "extrapolated": ninst < batch,
                   "sharding": "one process per host core"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{ninst} instances of the batch over {cores} processes, {args.steps} SQP iterations each (the rate "
                                   "of the full batch is the same: instances are independent); CPU restatement (SciPy/SuperLU "
                                   "interior point), not Ipopt"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "qp_solves": int(sum(o[1] for o in out)),
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="batch118")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--rounds", type=int, default=8, help="SQP rounds recorded for replay")
    ap.add_argument("--init-mu", dest="init_mu", type=float, default=1e5)
    ap.add_argument("--sqp-max-iter", dest="sqp_max_iter", type=int, default=100)
    ap.add_argument("--cpu-sample", type=int, default=12, help="QP subproblems solved by the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-spmv", action="store_true", help="skip the 2000-bus SpMV roofline leg")
    ap.add_argument("--no-device-eval", action="store_true", help="skip the full solve with the device-side evaluator")
    ap.add_argument("--spmv-batch", type=int, default=2048)
    ap.add_argument("--layout-G", dest="layout_G", type=int, default=0, help="instances interleaved per CTA (0 auto, 1 off, 2/4/8)")
    ap.add_argument("--layout-threads", dest="layout_threads", type=int, default=0)
    ap.add_argument("--layout-ctas", dest="layout_ctas", type=int, default=0)
    ap.add_argument("--layout-tail", dest="layout_tail", type=int, default=-1)
    ap.add_argument("--layout-handoff", dest="layout_handoff", type=int, default=None,
                    help="iteration quota before the resident launch takes an instance over (-1 auto, 0 off)")
    ap.add_argument("--no-single2000", action="store_true", help="skip the single-instance 2000-bus leg (BASELINE configs[3])")
    ap.add_argument("--groups", type=int, default=0,
                    help="independent groups of instances per GPU, each with its own engine handle / stream / host thread "
                         "(0 = auto: 4 from 768 instances per GPU, else 2; 1 = lock-step batch)")
    ap.add_argument("--repeats", type=int, default=5, help="the K-step timed region is run this many times; the MEDIAN is reported")
    ap.add_argument("--sequential-phases", dest="sequential_phases", action="store_true",
                    help="A/B: QP-phase and restoration-phase launches of a round one after the other (two calls) instead of side by side")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from sqpsolver_jl_b200 import capi
    from sqpsolver_jl_b200.host.batch import gather_results, pack_results, shard_range
    from sqpsolver_jl_b200.host.sqp_trust_region import GroupedBatchSqpTR, Parameters
    from sqpsolver_jl_b200.nlp.acopf import AcopfPolar

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line (the JSON): libraries that print there (NCCL's version banner) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    capi.build()
    net, batch, wname = make_workload(args)
    lo, hi = shard_range(batch, rank, world)
    Bl = hi - lo
    pd_all, qd_all = net.perturbed_loads(hi)  # Philox stream per instance id: only [lo,hi) is used
    pd, qd = pd_all[lo:hi], qd_all[lo:hi]
    nlp = AcopfPolar(net, pd=pd, qd=qd)
    kw = sqp_params(args)

    # ---- untimed set-up: run the real batched SQP on the device, record rounds sampled over the whole solve ----
    # The shard is driven as G independent groups of instances (GroupedBatchSqpTR: one engine handle, stream and host thread per
    # group): the launches of the other group fill the straggler tail of a launch and the host work between rounds.
    # measured (profiles/r02_tuning.md section 10): 1024 instances per GPU 71.3 / 60.2 / 57.5 ms per step with 1 / 2 / 4 groups
    # (cont.) finer groups keep helping until a group has about 32 instances: 256 per GPU 28.7 / 22.3 / 22.0 / 19.2 ms with
    # 1 / 2 / 4 / 8 groups, 128 per GPU 18.7 / 18.4 / 17.5 / 17.8 / 26.9 with 1 / 2 / 4 / 8 / 16, 512 per GPU 35.9 / 33.7 with 2 / 4
    G = args.groups if args.groups > 0 else (4 if Bl >= 512 else max(2, min(8, Bl // 32)))
    # every group has a host thread (the host-buffer C ABI blocks): keep them within the cores of the node
    if args.groups <= 0:
        G = min(G, max(2, (os.cpu_count() or 16) // max(1, world)))
    G = max(1, min(G, Bl))
    layout = dict(G=args.layout_G, threads=args.layout_threads, ctas_per_sm=args.layout_ctas, tail=args.layout_tail)
    if args.layout_handoff is not None:
        layout["handoff"] = args.layout_handoff
    sqp = GroupedBatchSqpTR(nlp, Bl, Parameters(**kw), groups=G, device=local, layout=layout)
    G = len(sqp.subs)
    gb = sqp.bounds
    # replayed rounds are sampled UNIFORMLY over the whole solve (round 1, 1 + T/R, ...): late rounds carry the stragglers
    stride = max(1, args.sqp_max_iter // max(1, args.rounds))
    sample_rounds = {1 + k * stride for k in range(args.rounds)}
    rec = [[] for _ in range(G)]  # per group: the recorded rounds (inputs of the hot path at the moment of the solve call)

    def make_hook(g, sub):
        orig_solve = sub.optimizer._solve
        Bg = gb[g][1] - gb[g][0]

        def snap(qp, fr):
            return {"round": sub.rounds, "dE": sub.dE.copy(), "h_val": sub.h_val.copy(), "df": sub.df.copy(), "E": sub.E.copy(),
                    "x": sub.x.copy(), "Delta": sub.Delta.copy(), "qp": qp, "fr": fr, "lam": sub.lam.copy(),
                    "mxU": sub.mult_x_U.copy(), "mxL": sub.mult_x_L.copy(), "mu": sub.mu.copy(), "f": sub.f.copy()}

        def solve_hook(phase, x_k, delta, E_override=None, active=None):
            if sub.rounds in sample_rounds:
                if phase == capi.PHASE_MIXED:  # both phases of the round in one call
                    rec[g].append(snap(np.asarray(active[0], np.int32).copy(), np.asarray(active[1], np.int32).copy()))
                elif phase in (capi.PHASE_QP, capi.PHASE_FR):
                    act = (np.ones(Bg, bool) if active is None else np.asarray(active, bool)).astype(np.int32)
                    if rec[g] and rec[g][-1]["round"] == sub.rounds:
                        rec[g][-1]["fr" if phase == capi.PHASE_FR else "qp"] = act
                    else:
                        z = np.zeros(Bg, np.int32)
                        rec[g].append(snap(act if phase == capi.PHASE_QP else z, act if phase == capi.PHASE_FR else z))
            return orig_solve(phase, x_k, delta, E_override, active)

        sub.optimizer._solve = solve_hook

    for g, sub in enumerate(sqp.subs):
        sub.mixed_phases = not args.sequential_phases
        make_hook(g, sub)
    t0 = time.perf_counter()
    sqp.run()
    t_sqp = time.perf_counter() - t0
    full = {"wall_s": t_sqp, "groups": G, "rounds": int(sqp.rounds), "sqp_iterations": int(sqp.iter.sum() - Bl + (sqp.ret != -1).sum()),
            "qp_solves": int(sqp.n_qp.sum()), "status_counts": {int(k): int(v) for k, v in zip(*np.unique(sqp.status, return_counts=True))},
            "device_s": sqp.timers["device"], "callbacks_s": sqp.timers["callbacks"],
            "solve_kernel_s": sqp.stats["solve_ms"] / 1e3}
    full["qp_solves_per_sec_wall"] = full["qp_solves"] / t_sqp                      # host NLP callbacks + PCIe + kernels
    full["qp_solves_per_sec_kernel"] = full["qp_solves"] / max(full["solve_kernel_s"], 1e-9)
    full["ms_per_round_kernel"] = 1e3 * full["solve_kernel_s"] / max(1, full["rounds"])
    full["converged_instances"] = int((sqp.status == 0).sum())
    full["note"] = ("device_s / callbacks_s / solve_kernel_s are summed over the %d groups, whose launches and host work overlap: they do "
                    "not add up to wall_s" % G) if G > 1 else ""
    res_local = pack_results(sqp.status, sqp.iter, sqp.obj_val)
    res_all = gather_results(res_local, batch, device=dev)  # the one collective of the path (NCCL, 16 B/instance)
    engs = [sub.optimizer.engine for sub in sqp.subs]
    eng = engs[0]
    R = min(len(r) for r in rec)
    assert R > 0

    # device-resident copies of the recorded inputs (for `value`)
    drec = [[{k: torch.from_numpy(np.ascontiguousarray(r[k])).to(dev) for k in ("dE", "h_val", "df", "E", "x", "Delta", "qp", "fr")}
             for r in rec[g][:R]] for g in range(G)]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    streams = [torch.cuda.ExternalStream(e.stream, device=dev) for e in engs]
    ptr = lambda t: t.data_ptr()

    def step_device(g, j):
        e, d, r = engs[g], drec[g][j % R], rec[g][j % R]
        if j % R == 0:
            e.set_options(warm_start=0)  # round 1 of a solve is cold
        e.update_nlp_device(ptr(d["dE"]), ptr(d["h_val"]), ptr(d["df"]), ptr(d["E"]))
        nq, nf = int(r["qp"].sum()), int(r["fr"].sum())
        if nq and nf and not args.sequential_phases:  # instances in both phases: one call, the two launches side by side
            e.solve_tr_mixed_device(ptr(d["x"]), ptr(d["Delta"]), ptr(d["qp"]), ptr(d["fr"]))
        else:
            if nq:
                e.solve_tr_device(capi.PHASE_QP, ptr(d["x"]), ptr(d["Delta"]), None, ptr(d["qp"]))
            if nf:
                e.solve_tr_device(capi.PHASE_FR, ptr(d["x"]), ptr(d["Delta"]), None, ptr(d["fr"]))
        if j % R == 0:
            e.set_options(warm_start=1)

    p_zero = [np.zeros_like(rec[g][0]["x"]) for g in range(G)]

    def step_host(g, j):
        e, r = engs[g], rec[g][j % R]
        if j % R == 0:
            e.set_options(warm_start=0)
        e.update_nlp(r["dE"], r["h_val"], r["df"], r["E"])
        e.merit(r["x"], p_zero[g], r["E"], r["f"], r["mu"])
        e.kt_residuals(r["lam"], r["mxU"], r["mxL"])
        out = None
        nq, nf = int(r["qp"].sum()), int(r["fr"].sum())
        if nq and nf and not args.sequential_phases:
            out = e.solve_tr_mixed(r["x"], r["Delta"], r["qp"], r["fr"])
        else:
            if nq:
                out = e.solve_tr(capi.PHASE_QP, r["x"], r["Delta"], active=r["qp"])
            if nf:
                out = e.solve_tr(capi.PHASE_FR, r["x"], r["Delta"], active=r["fr"])
        if j % R == 0:
            e.set_options(warm_start=1)
        return out

    def barrier():
        torch.cuda.synchronize()
        for e in engs:
            e.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def units_of(steps, warmup):
        return sum(int(rec[g][(warmup + j) % R]["qp"].sum() + rec[g][(warmup + j) % R]["fr"].sum()) for j in range(steps) for g in range(G))

    def region_ms(starts, ends):
        # device time of the region: from the first start event to the last end event (event timestamps are device-global)
        return max(s_.elapsed_time(e_) for s_ in starts for e_ in ends)

    def timed_device(steps, warmup, run_warmup=True):
        """K steps of every group enqueued back to back on the groups' streams (device-pointer API: nothing blocks), so that a
        group's next launch starts as soon as ITS previous one has drained -- the pipelining the grouped driver produces."""
        for j in range(warmup if run_warmup else 0):
            with torch.cuda.stream(streams[0]):
                flush.fill_(0xFF)  # also loads the fill kernel: its first launch (lazy module load) must not fall into the region
            for g in range(G):
                step_device(g, j)
        barrier()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(G)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(G)]
        for g in range(G):
            starts[g].record(streams[g])
        for j in range(steps):
            with torch.cuda.stream(streams[0]):
                flush.fill_(j & 0xFF)  # 256 MiB written between the steps of group 0 (the per-step working set exceeds L2 anyway)
            for g in range(G):
                step_device(g, warmup + j)
        for g in range(G):
            ends[g].record(streams[g])
        for e_ in ends:
            e_.synchronize()
        t = region_ms(starts, ends)
        barrier()
        return float(t), units_of(steps, warmup)

    def timed_host(steps, warmup, run_warmup=True):
        """The same steps through the blocking host-buffer C ABI, one host thread per group (as GroupedBatchSqpTR runs them)."""
        import threading
        for j in range(warmup if run_warmup else 0):
            for g in range(G):
                step_host(g, j)
        barrier()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(G)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(G)]
        for g in range(G):
            starts[g].record(streams[g])
        errs = []

        def work(g):
            try:
                for j in range(steps):
                    step_host(g, warmup + j)
            except BaseException as ex:  # noqa: BLE001
                errs.append(ex)

        th = [threading.Thread(target=work, args=(g,)) for g in range(G)]
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()
        if errs:
            raise errs[0]
        for g in range(G):
            ends[g].record(streams[g])
        for e_ in ends:
            e_.synchronize()
        t = region_ms(starts, ends)
        barrier()
        return float(t), units_of(steps, warmup)

    sampler = ClockSampler(local)
    sampler.start()
    l0 = sum(e.launch_count for e in engs)
    dev_regions = [timed_device(args.steps, args.warmup, k == 0) for k in range(max(1, args.repeats))]
    t_ms, units = float(np.median([t for t, _ in dev_regions])), dev_regions[0][1]
    launches = (sum(e.launch_count for e in engs) - l0) // max(1, args.repeats)
    clocks = sampler.stop()
    # untimed pass over the same steps, one at a time: per-instance iteration / factorisation counts (bit-reproducible) for the
    # bytes model, and the CUDA-event time of every solve launch run ALONE
    infos, solve_ms = [], []
    for j in range(args.steps):
        for g in range(G):
            ms0 = engs[g].solve_ms_total
            step_device(g, args.warmup + j)
            engs[g].sync()
            r = rec[g][(args.warmup + j) % R]
            infos.append((engs[g].fetch_info(), (r["qp"] | r["fr"]).astype(bool)))
            solve_ms.append(engs[g].solve_ms_total - ms0)
    for g in range(G):
        e = engs[g]
        e.reuse_outputs = True  # the host owns one set of result buffers and hands them to every call (no 60 MB allocation per step)
        # ... and page-locks its persistent arrays once (sqpqp_host_register), as the reference's host would its sqp.dE / h_val / df /
        # E / x / lambda vectors: the calls then copy straight between those arrays and the device (no staging memcpy)
        e.register_outputs = True
        for r in rec[g][:R]:
            e.register_host(*[r[k] for k in ("dE", "h_val", "df", "E", "x", "Delta", "lam", "mxU", "mxL", "mu", "f", "qp", "fr")])
        e.register_host(p_zero[g])
    host_regions = [timed_host(args.steps, args.warmup, k == 0) for k in range(max(1, args.repeats))]
    e_ms, e_units = float(np.median([t for t, _ in host_regions])), host_regions[0][1]

    # max over ranks of the time, sum over ranks of the units
    def allred(v, op):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    t_max = allred(t_ms, dist.ReduceOp.MAX) if world > 1 else t_ms
    e_max = allred(e_ms, dist.ReduceOp.MAX) if world > 1 else e_ms
    units_all = allred(float(units), dist.ReduceOp.SUM) if world > 1 else units
    e_units_all = allred(float(e_units), dist.ReduceOp.SUM) if world > 1 else e_units
    launches_all = allred(float(launches), dist.ReduceOp.SUM) if world > 1 else launches

    if rank == 0:
        n, m = nlp.n, nlp.m
        rpJ, ciJ, _ = eng.get_csr(0)
        rpH, ciH, _ = eng.get_csr(2)
        nnzJ, nnzH = int(ciJ.shape[0]), int(ciH.shape[0])
        chol = eng.chol_stats()
        alg = [algorithmic_bytes(i, sel, n, m, nnzJ, nnzH, chol) for i, sel in infos]
        ipm_it = float(np.sum([i["ipm_iters"][sel].sum() for i, sel in infos]) / max(1, np.sum([sel.sum() for i, sel in infos]))) if infos else 0.0
        ipm_max = int(max([i["ipm_iters"][sel].max() for i, sel in infos if sel.any()])) if infos else 0
        fallbacks = int(sum([(i["admm_iters"][sel] > 0).sum() for i, sel in infos]))
        # algorithmic bytes of EVERY solve launch of the timed region over the device time of the region (the launches of the
        # groups overlap, so a per-launch duration is not defined inside it); solve_ms: the same launches run one at a time
        ach = [float(np.sum(alg)) / (t_ms * 1e-3) / 1e9] if t_ms > 0 else []
        ach_alone = float(np.sum(alg)) / (float(np.sum(solve_ms)) * 1e-3) / 1e9 if np.sum(solve_ms) > 0 else None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = float(np.mean(ach)) if ach else 0.0
        kernel_name = eng.last_solve_kernel
        # dram__bytes of ONE launch of the dominant kernel from an `ncu --set full` capture of THIS bench command,
        # written next to the profile summary by tools/ncu_traffic.py (kernel name, launch shape and git revision inside);
        # reported only when it was taken for the kernel this run launched, else null
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            # (a launch covers one group of the shard)
            if kernel_name.startswith(tr.get("kernel", "?")) and tr.get("batch_per_gpu") == Bl // G and tr.get("workload") == args.workload:
                traffic = tr
        except (OSError, ValueError):
            pass
        b_it, b_f = bytes_model_ipm(n, m, nnzJ, nnzH, max(chol["nnzL"], 1))
        h2d = int(sum(sum(rec[g][0][k].nbytes for k in ("dE", "h_val", "df", "E", "x", "Delta", "qp")) + rec[g][0]["x"].nbytes * 2
                      + rec[g][0]["E"].nbytes * 2 + rec[g][0]["lam"].nbytes + 2 * rec[g][0]["mxU"].nbytes for g in range(G)))
        d2h = int(Bl * (3 * n + m + max(eng.S, 1)) * 8 + Bl * capi.INFO_DTYPE.itemsize + Bl * 8 * 6)
        line = {
            "metric": METRIC, "value": units_all / (t_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_max / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wname, "network": {"nbus": net.nbus, "nbranch": net.nbranch, "ngen": net.ngen, "seed": net.meta.get("seed")},
                       "n": n, "m": m, "nnzJ": nnzJ, "nnzH_sym": nnzH, "batch_total": batch, "batch_per_gpu": Bl, "sqp": kw,
                       "step": "one SQP iteration over the shard: COO value scatter + batched QP-subproblem solve kernel(s)",
                       "groups_per_gpu": G,
                       "pipelining": ("the shard is driven as %d independent groups of instances, each with its own engine handle and stream "
                                      "(host/sqp_trust_region.py: GroupedBatchSqpTR); the K timed steps of every group are enqueued back to "
                                      "back, so a group's next launch starts when ITS previous one has drained and fills the straggler "
                                      "tail of the other group's launch.  Every step's work completes inside the timed region" % G) if G > 1 else "none",
                       "replayed_rounds": R, "l2": "inputs larger than L2: consecutive steps replay DIFFERENT recorded rounds (%d input sets of %.0f MB "
                                                   "per GPU, cycled), and the per-instance work arrays (0.45 MB x batch_per_gpu) are rewritten by "
                                                   "every solve; in addition a 256 MiB buffer is written between the steps of one group"
                                                   % (R, h2d / 1e6),
                       "sharding": "contiguous instance blocks per rank, no data-path collective; one NCCL all-gather of 16 B/instance at the end"},
            "qp_solves_per_sec": units_all / (t_max * 1e-3),
            "timed_regions_ms": {"device": [round(t, 3) for t, _ in dev_regions], "host": [round(t, 3) for t, _ in host_regions],
                                 "note": "each region = exactly `steps` steps after the warm-up, bracketed by a barrier + synchronize; the "
                                         "MEDIAN region is reported (rank 0's regions shown)"},
            "e2e": {"value": e_units_all / (e_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e_max / args.steps,
                    "host_threads": G,
                    "note": "host buffers in, host buffers out through sqpqp_update_nlp / merit / kt_residuals / solve_tr(_mixed), one host thread "
                            "per group (the calls block); the caller's persistent "
                            "arrays (inputs and results) are page-locked once with sqpqp_host_register, so every step copies them "
                            "straight over the link (H2D and D2H inside the timed region, no staging memcpy)"},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
                         "traffic": traffic["dram_bytes"] if traffic else None,
                         "traffic_source": ({k: traffic[k] for k in ("profile", "launch", "algorithmic_bytes_same_launch", "git") if k in traffic}
                                            if traffic else None),
                         "kernel": kernel_name,  # from the library (sqpqp_last_solve_kernel): the launch rule lives there
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                         "achieved_one_launch_at_a_time": ach_alone,
                         "note": "achieved = algorithmic bytes of every solve launch of the timed region / device time of the region (launches of the "
                                 "groups overlap; achieved_one_launch_at_a_time: the same launches run alone, bytes / sum of their CUDA-event "
                                 "durations); bytes = per-instance fp64 values each "
                                 "phase of an interior-point iteration must touch once (%d B per iteration + %d B per Cholesky factorisation, "
                                 "DESIGN.md 5) x the device-counted iterations/factorisations of every instance in the launch + the shared int32 "
                                 "index programs once.  The kernel is bound by dependent-load latency of the level-scheduled sparse "
                                 "factorisation, not by HBM bandwidth (DESIGN.md 5.3)" % (b_it, b_f)},
            "solver": {"method": "interior point + batched sparse Cholesky (ADMM/PCG fallback)", "ipm_iters_mean": ipm_it,
                       "ipm_iters_max": ipm_max, "admm_fallbacks": fallbacks, "chol": chol},
            "full_sqp_solve": full,
            "note": ("value/e2e time the hot path (scatter + batched QP solve) on %d SQP rounds sampled uniformly over the solve; the NLP "
                     "callbacks are outside them.  The whole batched SQP solve incl. callbacks runs at full_sqp_solve.qp_solves_per_sec_wall. "
                     "%d of %d instances end with status 0: on this workload the reference algorithm as coded (Hessian multiplier sign, "
                     "sqp.jl:93) stalls at the iteration limit on BOTH sides (oracle and device, DESIGN.md section 8)"
                     % (R, full["converged_instances"], Bl)),
            "results_gathered": {"instances": int(res_all.shape[0]), "status_counts": {int(k): int(v) for k, v in zip(*np.unique(res_all["status"], return_counts=True))}},
        }
        if not args.no_spmv and world == 1:
            line["spmv"] = spmv_roofline(local, dev, peak, args.spmv_batch)
        if not args.no_single2000 and world == 1:
            line["single_instance_2000"] = single_instance_2000(local, peak)
        if not args.no_device_eval and world == 1:
            # the same full batched SQP solve with f, grad f, g, J and H values evaluated on the device (csrc/acopf.cuh,
            # SURVEY 8f rank 1) instead of by the host callbacks: only x and lambda go up per round
            sqp2 = GroupedBatchSqpTR(nlp, Bl, Parameters(**kw), groups=G, device=local, device_evaluator=True)
            for sub in sqp2.subs:
                sub.mixed_phases = not args.sequential_phases
            t0 = time.perf_counter()
            sqp2.run()
            line["full_sqp_solve_device_evaluator"] = {
                "wall_s": time.perf_counter() - t0, "groups": G, "rounds": int(sqp2.rounds), "qp_solves": int(sqp2.n_qp.sum()),
                "status_counts": {int(k): int(v) for k, v in zip(*np.unique(sqp2.status, return_counts=True))},
                "device_s": sqp2.timers["device"], "callbacks_s": sqp2.timers["callbacks"],
                "solve_kernel_s": sqp2.stats["solve_ms"] / 1e3,
                "max_rel_objective_diff_vs_host_evaluator": float(np.max(np.abs(sqp2.obj_val - sqp.obj_val) / np.maximum(1.0, np.abs(sqp.obj_val))))}
            sqp2.close()
        if not args.no_cpu_baseline and world == 1:
            # 1 thread and nproc threads, median of >= 5 repetitions after one warm-up, per QP solve
            import multiprocessing as mp
            jobs = []
            k = 0
            B0 = gb[0][1] - gb[0][0]  # subproblems of group 0 (instances gb[0][0] .. of the shard)
            while len(jobs) < max(6, args.cpu_sample + 1) and k < 64 * R:
                r = rec[0][k % R]
                b = (k // R) % B0
                if r["qp"][b]:
                    jobs.append((net, pd[gb[0][0] + b], qd[gb[0][0] + b], {"dE": r["dE"][b], "h_val": r["h_val"][b], "df": r["df"][b],
                                                     "E": r["E"][b], "x": r["x"][b], "Delta": float(r["Delta"][b])}))
                k += 1
            tcb0 = time.perf_counter()
            times = []
            for i, jb in enumerate(jobs):
                dt, st = _oracle_qp_worker(jb)
                if i > 0:  # the first solve is the warm-up
                    times.append(dt)
                if time.perf_counter() - tcb0 > 25.0 and len(times) >= 5:
                    break
            cores = os.cpu_count() or 1
            par = None
            try:
                with mp.Pool(cores) as pool:
                    pool.map(_oracle_qp_worker, [jobs[i % len(jobs)] for i in range(cores)])  # warm-up (imports, caches)
                    tp0 = time.perf_counter()
                    outp = pool.map(_oracle_qp_worker, [jobs[i % len(jobs)] for i in range(2 * cores)])
                    par = 2 * cores / (time.perf_counter() - tp0)
            except OSError:
                pass
            if times:
                line["cpu_baseline"] = {"value": 1.0 / float(np.median(times)), "unit": UNIT, "cores": 1, "kind": "port",
                                        "value_all_cores": par, "all_cores": cores,
                                        "median_s_per_qp": float(np.median(times)), "repetitions": len(times),
                                        "sample": f"{len(times)} of the replayed QP subproblems (rounds sampled over the solve) solved one after "
                                                  "another by the CPU oracle after one warm-up solve: 1 / median time; value_all_cores = "
                                                  f"{2 * cores} of them over a pool of {cores} processes.  CPU restatement (SciPy/SuperLU "
                                                  "interior point), not Ipopt"}
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
