"""Host-side mirror of the reference's SQP trust-region driver, running on the device engine.

In a deployment this file *is* ``src/algorithms/sqp_trust_region.jl`` (Julia, unchanged apart
from picking ``QpDevice`` as the sub-optimizer).  Julia is not available
in this image, so the same control flow is restated here in Python with the same names:

    SqpTR / run!                    sqp_trust_region.jl:6-223
    violation_of_linear_constraints :237-254
    sub_optimize_lp! / sub_optimize! / sub_optimize_soc!   :264-360
    compute_step!                   :370-380
    compute_qmodel / do_step!       :487-579
    eval_functions!, compute_phi, terminate_by_iterlimit   sqp.jl:86-117, 170-183, 215-224

What is different from the reference is WHERE the arithmetic runs: the COO scatter, the QP
solve, norm_violations, compute_qmodel, the merit value and KT_residuals are device calls
(``QpDevice`` / ``capi.Engine``); the host keeps only the NLP callbacks and the scalar
control flow.  The driver is batched: ``BatchSqpTR`` advances B independent instances that
share a sparsity pattern in lock-step rounds (one round = one ``while`` iteration of run! for
every unfinished instance), which is how the 1024-instance ACOPF workload is run; ``SqpTR``
is the B = 1 case.  Quirks of the reference are preserved as coded.
"""
from __future__ import annotations

import time

import numpy as np

from .parameters import Parameters
from .subproblem import QpDevice, OK_STATUSES, INFEASIBLE_STATUSES

RTOL_ISAPPROX = np.sqrt(np.finfo(float).eps)  # Julia's default isapprox rtol


def isapprox(a, b):
    return abs(a - b) <= RTOL_ISAPPROX * max(abs(a), abs(b))


def _ninf(v):
    return float(np.max(np.abs(v), initial=0.0))


class BatchSqpTR:
    def __init__(self, nlp, batch: int, params: Parameters | None = None, device: int = 0,
                 engine_options: dict | None = None, x0=None, device_evaluator: bool = False, layout: dict | None = None):
        self.problem = nlp
        self.options = params or Parameters()
        self.B = B = batch
        n, m = nlp.n, nlp.m
        x0 = nlp.x0 if x0 is None else x0
        self.x = np.ascontiguousarray(np.broadcast_to(np.asarray(x0, float), (B, n))).copy()
        self.p = np.zeros((B, n))
        self.p_soc = np.zeros((B, n))
        self.lam = np.zeros((B, m))
        self.mult_x_L = np.zeros((B, n))
        self.mult_x_U = np.zeros((B, n))
        self.p_lambda = np.zeros((B, m))
        self.p_mult_x_L = np.zeros((B, n))
        self.p_mult_x_U = np.zeros((B, n))
        self.f = np.full(B, np.nan)
        self.df = np.zeros((B, n))
        self.E = np.zeros((B, m))
        self.dE = np.zeros((B, nlp.nnz_jac_coo))
        self.h_val = np.zeros((B, nlp.nnz_hess_coo))
        self.phi = np.full(B, 1.0e20)
        self.mu = np.full(B, 1.0e4)
        self.Delta = np.full(B, 10.0)
        self.Delta_max = 1.0e8
        self.step_acceptance = np.ones(B, bool)
        self.prim_infeas = np.full(B, np.inf)
        self.dual_infeas = np.full(B, np.inf)
        self.sub_status = np.zeros(B, np.int32)
        self.feasibility_restoration = np.zeros(B, bool)
        self.iter = np.ones(B, np.int64)
        self.ret = np.full(B, -5, np.int64)
        self.done = np.zeros(B, bool)
        self.n_qp = np.zeros(B, np.int64)
        self.optimizer = QpDevice(nlp, batch=B, device=device, engine_options=engine_options, layout=layout)
        self.optimizer.create_model(None)
        # the persistent vectors of the SQP object (sqp.jl:16-59) are page-locked once, so that every call copies them straight
        # over the link (sqpqp_host_register); not worth it for a single small instance
        if B * max(n, m) >= (1 << 16):
            self.optimizer.engine.register_host(self.x, self.p, self.lam, self.mult_x_L, self.mult_x_U, self.df, self.E,
                                                self.dE, self.h_val)
            # ... and owns ONE set of solve results (every consumer below copies what it keeps before the next solve)
            self.optimizer.engine.reuse_outputs = True
            self.optimizer.engine.register_outputs = True
        # f, grad f, g and the J / H COO values evaluated on the device (csrc/acopf.cuh) instead of by the
        # host callbacks -- only x and lambda go up, only f, E, grad f come back
        self.device_evaluator = bool(device_evaluator)
        if self.device_evaluator:
            self.optimizer.engine.acopf_setup(nlp)
        self.mixed_phases = True  # a round with instances in both phases is ONE engine call (False: one call per phase)
        self.rounds = 0
        self.timers = {"callbacks": 0.0, "device": 0.0}
        self.trace = None
        self.trace_instances = None  # optional set of instance ids recorded in `trace` (None = all)

    # bounds may be per instance ([B,m]) or shared ([m])
    def _b(self, a, b):
        return a[b] if a.ndim == 2 else a

    # ------------------------------------------------------------------ callbacks
    def _eval_functions(self, mask):
        """eval_functions! (sqp.jl:86-104) for the masked instances + device scatter."""
        pr = self.problem
        t0 = time.perf_counter()
        idx = np.nonzero(mask)[0]
        if self.device_evaluator:
            if idx.size:
                f, E, df = self.optimizer.engine.acopf_eval_update(self.x, self.lam, np.asarray(mask, np.int32))
                self.f[idx], self.E[idx], self.df[idx] = f[idx], E[idx], df[idx]
            self.timers["device"] += time.perf_counter() - t0
            return
        if idx.size:
            xs = self.x[idx]
            f = np.atleast_1d(pr.eval_f(xs))
            df = np.empty((idx.size, pr.n)); pr.eval_grad_f(xs, df)
            E = np.empty((idx.size, pr.m)); self._eval_g(xs, E, idx)
            dE = np.empty((idx.size, pr.nnz_jac_coo)); pr.eval_jac_g(xs, dE)
            hv = np.empty((idx.size, pr.nnz_hess_coo)); pr.eval_h(xs, 1.0, self.lam[idx], hv)
            self.f[idx], self.df[idx], self.E[idx], self.dE[idx], self.h_val[idx] = f, df, E, dE, hv
        self.timers["callbacks"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        self.optimizer.update(self.dE, self.h_val, self.df, self.E)
        self.timers["device"] += time.perf_counter() - t0

    def _eval_g(self, xs, out, idx):
        """eval_g for a subset of instances (per-instance loads live inside the NLP object)."""
        pr = self.problem
        if getattr(pr, "batched_data", False):
            pr.eval_g_subset(xs, out, idx)
        else:
            pr.eval_g(xs, out)

    # ---------------------------------------------------------------------- run!
    def run(self, log=None):
        opt, pr, B = self.options, self.problem, self.B
        eng = self.optimizer.engine
        self.mu[:] = opt.init_mu
        self.Delta[:] = opt.tr_size
        t_start = time.perf_counter()

        # violation_of_linear_constraints + sub_optimize_lp! (sqp_trust_region.jl:112-122, 237-304)
        self.f = np.atleast_1d(np.asarray(pr.eval_f(self.x), float)).copy()
        nanf = np.isnan(self.f)
        self._eval_g(self.x, self.E, np.arange(B))
        ml = pr.num_linear_constraints
        gL = np.broadcast_to(pr.g_L, (B, pr.m))
        gU = np.broadcast_to(pr.g_U, (B, pr.m))
        lpviol = (np.sum(np.maximum(0.0, gL[:, :ml] - self.E[:, :ml]), 1) - np.sum(np.minimum(0.0, gU[:, :ml] - self.E[:, :ml]), 1)
                  + np.sum(np.maximum(0.0, pr.x_L - self.x), 1) - np.sum(np.minimum(0.0, pr.x_U - self.x), 1))
        self.ret[nanf] = -13
        self.done |= nanf
        need_lp = (lpviol > opt.tol_infeas) & ~self.done
        if need_lp.any():
            pr.eval_grad_f(self.x, self.df)
            pr.eval_jac_g(self.x, self.dE)
            self.optimizer.update(self.dE, self.h_val, self.df, self.E)
            xs, lam, mxU, mxL, st = self.optimizer.sub_optimize_lp(self.x, active=need_lp.astype(np.int32))
            self.n_qp[need_lp] += 1
            lam[:, ml:] = 0.0
            for b in np.nonzero(need_lp)[0]:
                if st[b] not in OK_STATUSES:
                    xs[b] = 0.0; lam[b] = 0.0; mxU[b] = 0.0; mxL[b] = 0.0
                for dst, src in ((self.x, xs), (self.lam, lam), (self.mult_x_U, mxU), (self.mult_x_L, mxL)):
                    v = src[b].copy()
                    v[np.abs(v) < 1e-10] = 0.0  # dropzeros! (utils.jl:16-22)
                    dst[b] = v
                self.sub_status[b] = st[b]

        while not self.done.all():
            self.rounds += 1
            act = ~self.done
            # terminate_by_iterlimit (sqp.jl:215-224)
            for b in np.nonzero(act & (self.iter > opt.max_iter))[0]:
                self.ret[b] = 6 if self.prim_infeas[b] <= opt.tol_infeas else -1
                self.done[b] = True
            act = ~self.done
            if not act.any():
                break
            ev = act & self.step_acceptance
            if ev.any():
                self._eval_functions(ev)
                t0 = time.perf_counter()
                mer = eng.merit(self.x, np.zeros_like(self.x), self.E, self.f, self.mu)  # norm_violations(sqp) (p = 1)
                kt = eng.kt_residuals(self.lam, self.mult_x_U, self.mult_x_L)
                self.timers["device"] += time.perf_counter() - t0
                self.prim_infeas[ev] = mer["viol0"][ev]
                self.dual_infeas[ev] = kt[ev]
            # compute_step! (sqp_trust_region.jl:370-380)
            t0 = time.perf_counter()
            qp_mask = act & ~self.feasibility_restoration
            fr_mask = act & self.feasibility_restoration
            new_lam = np.zeros_like(self.lam); new_U = np.zeros_like(self.mult_x_U); new_L = np.zeros_like(self.mult_x_L)
            # instances in both phases: one call, the two launches side by side (sqpqp_solve_tr_mixed); else the phase's own call
            if self.mixed_phases and qp_mask.any() and fr_mask.any():
                calls = ((act, lambda x, d, active: self.optimizer.sub_optimize_mixed(x, d, qp_mask.astype(np.int32), fr_mask.astype(np.int32))),)
            else:
                calls = ((qp_mask, self.optimizer.sub_optimize), (fr_mask, self.optimizer.sub_optimize_FR))
            for mask, fn in calls:
                if mask.any():
                    p, lam, mxU, mxL, _, st = fn(self.x, self.Delta, active=mask.astype(np.int32))
                    self.p[mask] = p[mask]; new_lam[mask] = lam[mask]; new_U[mask] = mxU[mask]; new_L[mask] = mxL[mask]
                    self.sub_status[mask] = st[mask]
                    self.n_qp[mask] += 1
                    if self.trace is not None:
                        info = self.optimizer.last_info
                        for b in np.nonzero(mask)[0]:
                            if self.trace_instances is not None and int(b) not in self.trace_instances:
                                continue
                            self.trace.append({"b": int(b), "iter": int(self.iter[b]), "fr": bool(self.feasibility_restoration[b]),
                                               "x": self.x[b].copy(), "Delta": float(self.Delta[b]), "dE": self.dE[b].copy(),
                                               "h_val": self.h_val[b].copy(), "df": self.df[b].copy(), "E": self.E[b].copy(),
                                               "p": p[b].copy(), "lambda_qp": lam[b].copy(), "mult_x_U": mxU[b].copy(),
                                               "mult_x_L": mxL[b].copy(), "status": int(st[b]), "info": info[b].copy(),
                                               "lam": self.lam[b].copy()})
            self.timers["device"] += time.perf_counter() - t0
            self.p_lambda[act] = new_lam[act] - self.lam[act]
            self.p_mult_x_L[act] = new_L[act] - self.mult_x_L[act]
            self.p_mult_x_U[act] = new_U[act] - self.mult_x_U[act]
            mmax = np.maximum(np.max(np.abs(self.lam), axis=1, initial=0.0),
                              np.maximum(np.max(np.abs(self.mult_x_L), axis=1), np.max(np.abs(self.mult_x_U), axis=1)))
            self.mu[act] = np.maximum(self.mu[act], mmax[act])

            # status branches, phi, termination tests (sqp_trust_region.jl:144-204)
            step = np.zeros(B, bool)
            for b in np.nonzero(act)[0]:
                st = self.sub_status[b]
                pinf = _ninf(self.p[b])
                if st in OK_STATUSES:
                    if self.Delta[b] == self.Delta_max and isapprox(pinf, self.Delta[b]):
                        self.ret[b] = 4; self.done[b] = True; continue
                elif st in INFEASIBLE_STATUSES:
                    if self.feasibility_restoration[b]:
                        self.ret[b] = 6 if self.prim_infeas[b] <= opt.tol_infeas else 2
                        self.done[b] = True
                    else:
                        self.feasibility_restoration[b] = True
                        self._log(log, b)
                        self.iter[b] += 1
                    continue
                else:
                    if self.prim_infeas[b] <= opt.tol_infeas * 10.0:
                        self.ret[b] = 6
                    self.done[b] = True
                    continue
                if self.step_acceptance[b]:  # compute_phi(x, 0, p)
                    nv = self.prim_infeas[b]
                    self.phi[b] = nv if self.feasibility_restoration[b] else self.f[b] + self.mu[b] * nv
                self._log(log, b)
                if pinf <= opt.tol_direction:
                    if self.feasibility_restoration[b]:
                        self.feasibility_restoration[b] = False
                        self.iter[b] += 1
                    else:
                        self.ret[b] = 0; self.done[b] = True
                    continue
                if (self.prim_infeas[b] <= opt.tol_infeas and self.dual_infeas[b] <= opt.tol_residual
                        and not isapprox(self.Delta[b], pinf) and not self.feasibility_restoration[b]):
                    self.ret[b] = 0; self.done[b] = True
                    continue
                step[b] = True
            if step.any():
                self._do_step(step)
                for b in np.nonzero(step)[0]:
                    if self.feasibility_restoration[b] and self.step_acceptance[b]:
                        self.feasibility_restoration[b] = False
                    self.iter[b] += 1

        self.obj_val = np.atleast_1d(np.asarray(pr.eval_f(self.x), float))
        self.status = self.ret.copy()
        self.mult_g = -self.lam
        self.elapsed = time.perf_counter() - t_start
        return self

    # ------------------------------------------------------------------ do_step!
    def _do_step(self, step):
        opt, pr, B = self.options, self.problem, self.B
        eng = self.optimizer.engine
        idx = np.nonzero(step)[0]
        t0 = time.perf_counter()
        xt = self.x + self.p
        if self.device_evaluator and not opt.use_soc:
            # compute_phi(x, 1, p) (sqp.jl:170-183): f and g at the trial point evaluated on the device and left there for the
            # merit call -- neither array crosses the link
            eng.acopf_eval_trial(xt, step.astype(np.int32))
            mer = eng.merit(self.x, self.p, None, None, self.mu, self.feasibility_restoration.astype(np.int32))
            self.timers["device"] += time.perf_counter() - t0
            E_t = None
        else:
            f_t = self.f.copy()
            E_t = self.E.copy()
            f_t[idx] = np.atleast_1d(pr.eval_f(xt[idx]))
            Et = np.empty((idx.size, pr.m)); self._eval_g(xt[idx], Et, idx); E_t[idx] = Et
            self.timers["callbacks"] += time.perf_counter() - t0
            t0 = time.perf_counter()
            mer = eng.merit(self.x, self.p, E_t, f_t, self.mu, self.feasibility_restoration.astype(np.int32))
            self.timers["device"] += time.perf_counter() - t0
        soc_need = np.zeros(B, bool)
        ared = np.zeros(B); q0 = mer["q0"]
        for b in idx:
            phi_k = mer["phi_trial"][b]
            ared[b] = self.phi[b] - phi_k
            pred = 1.0 if self.feasibility_restoration[b] else q0[b] - mer["qk"][b]
            with np.errstate(divide="ignore", invalid="ignore"):
                rho = np.float64(ared[b]) / np.float64(pred)
            if ared[b] > 0 and rho > 0:
                self._accept(b, self.p[b])
            else:
                c_k = mer["viol_trial"][b]
                if opt.use_soc and c_k > 0 and not self.feasibility_restoration[b]:
                    soc_need[b] = True
                else:
                    self._reject(b)
        if soc_need.any():  # sub_optimize_soc! (sqp_trust_region.jl:341-360, 544-572)
            sidx = np.nonzero(soc_need)[0]
            t0 = time.perf_counter()
            Jp = self._jac_times(self.p)
            E_soc = E_t - Jp
            ps, _, _, _, _, st = self.optimizer.sub_optimize_soc(self.x, self.Delta, E_soc, active=soc_need.astype(np.int32))
            self.n_qp[soc_need] += 1
            self.p_soc[sidx] = self.p[sidx] + ps[sidx]
            xs = self.x + self.p_soc
            f_s = self.f.copy(); E_s = self.E.copy()
            f_s[sidx] = np.atleast_1d(pr.eval_f(xs[sidx]))
            Es_ = np.empty((sidx.size, pr.m)); self._eval_g(xs[sidx], Es_, sidx); E_s[sidx] = Es_
            mer2 = eng.merit(self.x, self.p_soc, E_s, f_s, self.mu, self.feasibility_restoration.astype(np.int32))
            self.timers["device"] += time.perf_counter() - t0
            for b in sidx:
                ared_s = self.phi[b] - mer2["phi_trial"][b]
                pred_s = q0[b] - mer2["qk"][b]
                with np.errstate(divide="ignore", invalid="ignore"):
                    rho_s = np.float64(ared_s) / np.float64(pred_s)
                if ared_s > 0 and rho_s > 0:
                    self.x[b] = self.x[b] + self.p_soc[b]
                    self.lam[b] += self.p_lambda[b]
                    self.mult_x_L[b] += self.p_mult_x_L[b]
                    self.mult_x_U[b] += self.p_mult_x_U[b]
                    self.step_acceptance[b] = True
                else:
                    self._reject(b)

    def _accept(self, b, p):
        self.x[b] = self.x[b] + p
        self.lam[b] += self.p_lambda[b]
        self.mult_x_L[b] += self.p_mult_x_L[b]
        self.mult_x_U[b] += self.p_mult_x_U[b]
        if isapprox(self.Delta[b], _ninf(p)):
            self.Delta[b] = min(2 * self.Delta[b], self.Delta_max)
        self.step_acceptance[b] = True

    def _reject(self, b):
        self.Delta[b] = max(0.5 * min(self.Delta[b], _ninf(self.p[b])), 0.1 * self.options.tol_direction)
        self.step_acceptance[b] = False

    def _jac_times(self, p):
        """``Jacobian * p`` per instance on device (sqp_trust_region.jl:343)."""
        return self.optimizer.engine.jac_times(p)

    def _log(self, log, b):
        if log is None:
            return
        log.append({"b": int(b), "iter": int(self.iter[b]), "fr": bool(self.feasibility_restoration[b]),
                    "accept": bool(self.step_acceptance[b]), "f": float(self.f[b]), "phi": float(self.phi[b]),
                    "mu": float(self.mu[b]), "Delta": float(self.Delta[b]), "pinf": _ninf(self.p[b]),
                    "inf_pr": float(self.prim_infeas[b]), "inf_du": float(self.dual_infeas[b]),
                    "sub_status": int(self.sub_status[b])})

    def close(self):
        self.optimizer.close()


class GroupedBatchSqpTR:
    """A batch driven as G independent GROUPS of instances, one :class:`BatchSqpTR` (its own engine handle and stream) and one
    host thread per group.

    The instances of a batch never interact, so nothing forces them through the SQP rounds in lock step.  With
    one group the GPU idles at the end of every solve launch while the few slow subproblems finish (15-24 % of a 1024-instance
    launch, profiles/r02_tuning.md section 8), and it idles completely while the host evaluates the NLP callbacks and the
    ratio test.  With G >= 2 groups the launches of the other groups fill both gaps: a group's next round is queued as soon
    as ITS slowest instance and ITS host work are done.  Same per-instance arithmetic and results as :class:`BatchSqpTR`
    (every group is one); only the order in which the hardware sees the work changes.

    ``nlp.subset(lo, hi)`` must return the NLP of instances lo..hi-1 (per-instance data sliced)."""

    def __init__(self, nlp, batch: int, params: Parameters | None = None, groups: int = 2, device: int = 0,
                 engine_options: dict | None = None, x0=None, device_evaluator: bool = False, layout: dict | None = None):
        self.problem, self.B = nlp, batch
        G = max(1, min(int(groups), batch))
        self.bounds = [(g * batch // G, (g + 1) * batch // G) for g in range(G)]
        self.subs = []
        if G > 1:  # the hand-off to the resident launch assumes that a shard has the GPU to itself (one wave): off for groups
            layout = dict(layout or {})
            layout.setdefault("handoff", 0)
        for lo, hi in self.bounds:
            x0g = None if x0 is None or np.ndim(x0) == 1 else np.asarray(x0)[lo:hi]
            sub_nlp = nlp.subset(lo, hi) if hasattr(nlp, "subset") else nlp  # an NLP without per-instance data is its own subset
            self.subs.append(BatchSqpTR(sub_nlp, hi - lo, params, device=device, engine_options=engine_options,
                                        x0=x0 if x0g is None else x0g, device_evaluator=device_evaluator, layout=layout))
        self.options = self.subs[0].options

    def run(self, log=None):
        import threading
        errs = []

        logs = [([] if log is not None else None) for _ in self.subs]

        def work(sub, lg):
            try:
                sub.run(lg)
            except BaseException as e:  # noqa: BLE001 -- re-raised in the caller's thread
                errs.append(e)

        t0 = time.perf_counter()
        threads = [threading.Thread(target=work, args=(s, lg)) for s, lg in zip(self.subs, logs)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errs:
            raise errs[0]
        self.elapsed = time.perf_counter() - t0
        if log is not None:  # instance ids of the whole batch
            for (lo, _), lg in zip(self.bounds, logs):
                for e in lg:
                    e["b"] += lo
                    log.append(e)
        cat = lambda name: np.concatenate([getattr(s, name) for s in self.subs])  # noqa: E731
        for name in ("x", "lam", "mult_x_L", "mult_x_U", "mult_g", "obj_val", "status", "ret", "n_qp", "iter", "prim_infeas",
                     "dual_infeas", "f"):
            setattr(self, name, cat(name))
        self.rounds = max(s.rounds for s in self.subs)
        self.timers = {k: sum(s.timers[k] for s in self.subs) for k in self.subs[0].timers}
        self.stats = {k: sum(s.optimizer.stats[k] for s in self.subs) for k in self.subs[0].optimizer.stats}
        return self

    def close(self):
        for s in self.subs:
            s.close()


class SqpTR:
    """Single-instance view of :class:`BatchSqpTR` (the reference's ``SqpTR``)."""

    def __init__(self, nlp, params: Parameters | None = None, device: int = 0, engine_options: dict | None = None):
        self.batch = BatchSqpTR(nlp, 1, params, device, engine_options)

    def run(self, log=None, trace=None):
        bt = self.batch
        bt.trace = trace
        bt.run(log)
        self.x = bt.x[0]
        self.status = int(bt.status[0])
        self.obj_val = float(bt.obj_val[0])
        self.iter = int(bt.iter[0])
        self.n_qp = int(bt.n_qp[0])
        self.lam = bt.lam[0]
        self.mult_x_L, self.mult_x_U = bt.mult_x_L[0], bt.mult_x_U[0]
        self.prim_infeas, self.dual_infeas = float(bt.prim_infeas[0]), float(bt.dual_infeas[0])
        self.elapsed = bt.elapsed
        self.stats = bt.optimizer.stats
        self.timers = bt.timers
        return self

    def close(self):
        self.batch.close()
