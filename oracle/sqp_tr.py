"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's SQP trust-region driver.

Line-by-line restatement (plain numpy, float64) of

  SqpTR ctor                     sqp_trust_region.jl:26-90
  run!                           sqp_trust_region.jl:98-223
  violation_of_linear_constraints  :237-254
  sub_optimize_lp!               :264-304      (+ dropzeros!, utils.jl:16-22)
  sub_optimize! / _soc!          :314-360
  compute_step!                  :370-380
  compute_qmodel                 :487-508
  do_step!                       :515-579
  eval_functions!/eval_Jacobian! sqp.jl:86-117
  compute_phi                    sqp.jl:170-183
  terminate_by_iterlimit         sqp.jl:215-224
  KT_residuals / norm_violations common.jl:14-23, 54-77

Quirks are preserved as coded: mu is updated from the
pre-step multipliers; ``sqp.ret == -3`` is a no-op; prim_infeas is the L1 norm
including variable-bound violations; the termination test needs
``Delta !~ |p|_inf``.  ``isapprox`` is Julia's default (rtol = sqrt(eps)).

The QP sub-solver is :class:`oracle.subproblem.QpOracle` (stand-in for Ipopt;
parity unpinned at the QP level -- see oracle/qp_solver.py).
"""
from __future__ import annotations

import time

import numpy as np

from . import qp_solver as qs
from .coo import CooMatrix, SymCooMatrix
from .subproblem import QpData, QpOracle, sub_optimize_lp

RTOL_ISAPPROX = np.sqrt(np.finfo(float).eps)


def isapprox(a, b):
    return abs(a - b) <= RTOL_ISAPPROX * max(abs(a), abs(b))


def norm_violations(E, g_L, g_U, x, x_L, x_U, p=1):
    """common.jl:54-77."""
    vc = np.where(E > g_U, E - g_U, np.where(E < g_L, g_L - E, 0.0))
    vx = np.where(x > x_U, x - x_U, np.where(x < x_L, x_L - x, 0.0))
    viol = np.concatenate([vc, vx])
    if p == 1:
        return float(np.sum(np.abs(viol)))
    return float(np.max(np.abs(viol), initial=0.0))


def KT_residuals(df, lam, mult_x_U, mult_x_L, J):
    """common.jl:14-23 (J is scipy CSR)."""
    res = np.max(np.abs(df + J.T @ lam + mult_x_U - mult_x_L), initial=0.0)
    scalar = max(1.0, np.max(np.abs(df), initial=0.0), np.max(np.abs(mult_x_U), initial=0.0),
                 np.max(np.abs(mult_x_L), initial=0.0))
    rown = np.sqrt(np.asarray(J.multiply(J).sum(axis=1)).ravel())
    if lam.shape[0]:
        scalar = max(scalar, float(np.max(np.abs(lam) * rown)))
    return res / scalar


class Parameters:
    """parameters.jl:1-30 (live fields only)."""

    def __init__(self, **kw):
        self.algorithm = "SQP-TR"
        self.OutputFlag = 0
        self.tol_direction = 1e-8
        self.tol_residual = 1e-8
        self.tol_infeas = 1e-8
        self.max_iter = 3000
        self.init_mu = 1.0
        self.tr_size = 10.0
        self.use_soc = False
        for k, v in kw.items():
            if not hasattr(self, k):
                raise KeyError(k)
            setattr(self, k, v)


class SqpTROracle:
    def __init__(self, nlp, params: Parameters | None = None, sub_factory=None, qp_tol=1e-10, trace=None):
        self.problem = nlp
        self.options = params or Parameters()
        n, m = nlp.n, nlp.m
        self.x = np.array(nlp.x0, float)
        self.p = np.zeros(n)
        self.p_soc = np.zeros(n)
        self.p_slack = {}
        self.lam = np.zeros(m)
        self.mult_x_L = np.zeros(n)
        self.mult_x_U = np.zeros(n)
        self.df = np.zeros(n)
        self.E = np.zeros(m)
        self.dE = np.zeros(nlp.nnz_jac_coo)
        self.h_val = np.zeros(nlp.nnz_hess_coo)
        self.Jacobian = CooMatrix(nlp.j_row, nlp.j_col, m, n)
        self.Hessian = SymCooMatrix(nlp.h_row, nlp.h_col, n)
        self.p_lambda = np.zeros(m)
        self.p_mult_x_L = np.zeros(n)
        self.p_mult_x_U = np.zeros(n)
        self.phi = 1.0e20
        self.mu = 1.0e4
        self.Delta = 10.0
        self.Delta_max = 1.0e8
        self.step_acceptance = True
        self.prim_infeas = np.inf
        self.dual_infeas = np.inf
        self.optimizer = None
        self.sub_status = None
        self.feasibility_restoration = False
        self.iter = 1
        self.ret = -5
        self.f = np.nan
        self.sub_factory = sub_factory or (lambda data: QpOracle(data, qp_tol=qp_tol))
        self.qp_tol = qp_tol
        self.trace = trace  # optional list receiving one dict per QP solve
        self.n_qp = 0
        self.qp_time = 0.0

    # ------------------------------------------------------------------ evals
    def _J(self):
        return self.Jacobian.to_scipy()

    def _H(self):
        return self.Hessian.to_scipy()

    def eval_Jacobian(self):
        self.problem.eval_jac_g(self.x, self.dE)
        self.Jacobian.fill(self.dE)

    def eval_functions(self):
        pr = self.problem
        self.f = float(pr.eval_f(self.x))
        pr.eval_grad_f(self.x, self.df)
        pr.eval_g(self.x, self.E)
        self.eval_Jacobian()
        pr.eval_h(self.x, 1.0, self.lam, self.h_val)
        self.Hessian.fill(self.h_val)

    def norm_viol_current(self):
        pr = self.problem
        return norm_violations(self.E, pr.g_L, pr.g_U, self.x, pr.x_L, pr.x_U, 1)

    def qp_data(self, b=None):
        pr = self.problem
        return QpData(self._H(), self.df, self._J(), self.E if b is None else b, pr.g_L, pr.g_U, pr.x_L, pr.x_U,
                      pr.num_linear_constraints)

    def compute_phi(self, x, alpha, p):
        pr = self.problem
        tmpx = x + alpha * p
        f = self.f
        tmpE = self.E.copy()
        if alpha > 0.0:
            f = float(pr.eval_f(tmpx))
            pr.eval_g(tmpx, tmpE)
        nv = norm_violations(tmpE, pr.g_L, pr.g_U, tmpx, pr.x_L, pr.x_U, 1)
        return nv if self.feasibility_restoration else f + self.mu * nv

    def compute_qmodel(self, p, with_step):
        pr = self.problem
        qval = 0.0
        if with_step:
            qval += self.df @ p + 0.5 * (p @ (self._H() @ p))
            tmpx = self.x + p
            tmpE = self.E + self._J() @ p
        else:
            tmpx, tmpE = self.x, self.E
        return qval + self.mu * norm_violations(tmpE, pr.g_L, pr.g_U, tmpx, pr.x_L, pr.x_U, 1)

    # ---------------------------------------------------------------- sub-solves
    def violation_of_linear_constraints(self, x):
        pr = self.problem
        self.f = float(pr.eval_f(self.x))
        if not np.isnan(self.f):
            pr.eval_g(x, self.E)
        ml = pr.num_linear_constraints
        lpviol = np.sum(np.maximum(0.0, pr.g_L[:ml] - self.E[:ml])) - np.sum(np.minimum(0.0, pr.g_U[:ml] - self.E[:ml]))
        lpviol += np.sum(np.maximum(0.0, pr.x_L - x)) - np.sum(np.minimum(0.0, pr.x_U - x))
        return float(lpviol)

    def sub_optimize_lp(self):
        pr = self.problem
        self.f = float(pr.eval_f(self.x))
        pr.eval_grad_f(self.x, self.df)
        self.eval_Jacobian()
        t0 = time.perf_counter()
        self.x, self.lam, self.mult_x_U, self.mult_x_L, self.sub_status = sub_optimize_lp(
            self._J(), pr.g_L, pr.g_U, pr.x_L, pr.x_U, self.x, pr.num_linear_constraints, pr.m, qp_tol=self.qp_tol)
        self.qp_time += time.perf_counter() - t0
        self.n_qp += 1
        for v in (self.x, self.lam, self.mult_x_U, self.mult_x_L):
            v[np.abs(v) < 1e-10] = 0.0  # dropzeros!

    def sub_optimize(self):
        if self.optimizer is None:
            self.optimizer = self.sub_factory(self.qp_data())
            self.optimizer.create_model(self.Delta)
        else:
            self.optimizer.data = self.qp_data()
        t0 = time.perf_counter()
        if self.feasibility_restoration:
            out = self.optimizer.sub_optimize_FR(self.x, self.Delta)
        else:
            out = self.optimizer.sub_optimize(self.x, self.Delta)
        self.qp_time += time.perf_counter() - t0
        self.n_qp += 1
        if self.trace is not None:
            self.trace.append({
                "iter": self.iter, "fr": self.feasibility_restoration, "x": self.x.copy(), "Delta": self.Delta,
                "lam": self.lam.copy(), "dE": self.dE.copy(), "h_val": self.h_val.copy(), "df": self.df.copy(),
                "E": self.E.copy(), "p": out[0].copy(), "lambda_qp": out[1].copy(), "mult_x_U": out[2].copy(),
                "mult_x_L": out[3].copy(), "status": out[5],
            })
        return out

    def sub_optimize_soc(self):
        pr = self.problem
        E_soc = np.zeros(pr.m)
        pr.eval_g(self.x + self.p, E_soc)
        E_soc = E_soc - self._J() @ self.p
        self.optimizer.data = self.qp_data(b=E_soc)
        t0 = time.perf_counter()
        p, *_ = self.optimizer.sub_optimize(self.x, self.Delta)
        self.qp_time += time.perf_counter() - t0
        self.n_qp += 1
        self.p_soc = self.p + p

    def compute_step(self):
        self.p, lam, mult_x_U, mult_x_L, self.p_slack, self.sub_status = self.sub_optimize()
        self.p_lambda = lam - self.lam
        self.p_mult_x_L = mult_x_L - self.mult_x_L
        self.p_mult_x_U = mult_x_U - self.mult_x_U
        ninf = lambda v: float(np.max(np.abs(v), initial=0.0))
        self.mu = max(self.mu, ninf(self.lam), ninf(self.mult_x_L), ninf(self.mult_x_U))

    # -------------------------------------------------------------------- do_step!
    def do_step(self):
        opt = self.options
        phi_k = self.compute_phi(self.x, 1.0, self.p)
        ared = self.phi - phi_k
        pred = 1.0
        q_0 = None
        if not self.feasibility_restoration:
            q_0 = self.compute_qmodel(self.p, False)
            q_k = self.compute_qmodel(self.p, True)
            pred = q_0 - q_k
        with np.errstate(divide="ignore", invalid="ignore"):
            rho = np.float64(ared) / np.float64(pred)
        pinf = float(np.max(np.abs(self.p), initial=0.0))
        if ared > 0 and rho > 0:
            self.x = self.x + self.p
            self.lam = self.lam + self.p_lambda
            self.mult_x_L = self.mult_x_L + self.p_mult_x_L
            self.mult_x_U = self.mult_x_U + self.p_mult_x_U
            if isapprox(self.Delta, pinf):
                self.Delta = min(2 * self.Delta, self.Delta_max)
            self.step_acceptance = True
        else:
            pr = self.problem
            tmpx = self.x + self.p
            tmpE = np.zeros(pr.m)
            pr.eval_g(tmpx, tmpE)
            c_k = norm_violations(tmpE, pr.g_L, pr.g_U, tmpx, pr.x_L, pr.x_U, 1)
            perform_soc = False
            if opt.use_soc and c_k > 0 and not self.feasibility_restoration:
                self.sub_optimize_soc()
                phi_soc = self.compute_phi(self.x, 1.0, self.p_soc)
                ared = self.phi - phi_soc
                q_soc = self.compute_qmodel(self.p_soc, True)
                pred = q_0 - q_soc
                with np.errstate(divide="ignore", invalid="ignore"):
                    rho_soc = np.float64(ared) / np.float64(pred)
                if ared > 0 and rho_soc > 0:
                    self.x = self.x + self.p_soc
                    self.lam = self.lam + self.p_lambda
                    self.mult_x_L = self.mult_x_L + self.p_mult_x_L
                    self.mult_x_U = self.mult_x_U + self.p_mult_x_U
                    self.step_acceptance = True
                    perform_soc = True
            if not perform_soc:
                self.Delta = max(0.5 * min(self.Delta, pinf), 0.1 * opt.tol_direction)
                self.step_acceptance = False

    # ------------------------------------------------------------------------ run!
    def run(self, log=None):
        opt, pr = self.options, self.problem
        self.mu = opt.init_mu
        self.Delta = opt.tr_size
        t_start = time.perf_counter()
        lpviol = self.violation_of_linear_constraints(self.x)
        if np.isnan(self.f):
            self.status = -13
            return self
        elif lpviol > opt.tol_infeas:
            self.sub_optimize_lp()
        while True:
            if self.iter > opt.max_iter:
                self.ret = -1
                if self.prim_infeas <= opt.tol_infeas:
                    self.ret = 6
                break
            if self.step_acceptance:
                self.eval_functions()
                self.prim_infeas = self.norm_viol_current()
                self.dual_infeas = KT_residuals(self.df, self.lam, self.mult_x_U, self.mult_x_L, self._J())
            self.compute_step()
            if self.sub_status in qs.OK_STATUSES:
                if self.Delta == self.Delta_max and isapprox(float(np.max(np.abs(self.p), initial=0.0)), self.Delta):
                    self.ret = 4
                    break
            elif self.sub_status in qs.INFEASIBLE_STATUSES:
                if self.feasibility_restoration:
                    self.ret = 6 if self.prim_infeas <= opt.tol_infeas else 2
                    break
                else:
                    self.feasibility_restoration = True
                    self._log(log)
                    self.iter += 1
                    continue
            else:
                if self.prim_infeas <= opt.tol_infeas * 10.0:
                    self.ret = 6
                break
            if self.step_acceptance:
                self.phi = self.compute_phi(self.x, 0.0, self.p)
            self._log(log)
            pinf = float(np.max(np.abs(self.p), initial=0.0))
            if pinf <= opt.tol_direction:
                if self.feasibility_restoration:
                    self.feasibility_restoration = False
                    self.iter += 1
                    continue
                else:
                    self.ret = 0
                    break
            if (self.prim_infeas <= opt.tol_infeas and self.dual_infeas <= opt.tol_residual
                    and not isapprox(self.Delta, pinf) and not self.feasibility_restoration):
                self.ret = 0
                break
            self.do_step()
            if self.feasibility_restoration and self.step_acceptance:
                self.feasibility_restoration = False
            self.iter += 1
        self.obj_val = float(pr.eval_f(self.x))
        self.status = int(self.ret)
        self.mult_g = -self.lam
        self.elapsed = time.perf_counter() - t_start
        return self

    def _log(self, log):
        if log is None:
            return
        log.append({
            "iter": self.iter, "fr": self.feasibility_restoration, "accept": self.step_acceptance, "f": self.f,
            "phi": self.phi, "mu": self.mu, "Delta": self.Delta, "pinf": float(np.max(np.abs(self.p), initial=0.0)),
            "inf_pr": self.prim_infeas, "inf_du": self.dual_infeas, "sub_status": self.sub_status,
        })
