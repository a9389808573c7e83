"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's SQP line-search driver.

Restates (plain numpy, float64)

  SqpLS ctor / run!           sqp_line_search.jl:4-64, 71-251
  compute_mu_rule2!           sqp_line_search.jl:280-291   (compute_mu! = rule 2, :269)
  compute_alpha (Armijo)      sqp_line_search.jl:303-334
  compute_phi                 sqp.jl:170-183
  compute_derivative          sqp.jl:190-213 + merit.jl:13-17
  norm_complementarity        common.jl:30-47
  KT_residuals / norm_violations  common.jl:14-23, 54-77

The reference file is NOT compiled by the reference (`# include("sqp_line_search.jl")`, sqp.jl:226) and is stale: it
calls a three-argument ``sub_optimize!(sqp, model, 1000.0)`` that exists nowhere, and it multiplies the per-row
penalty VECTOR ``sqp.μ`` with the scalar L1 violation in ``compute_phi`` and then compares the result with a scalar.
There is therefore no as-coded behaviour to pin; this restatement fixes the two gaps in the only way that keeps every
other line as written, and says so:

  * the QP is the trust-region subproblem with the literal Delta = 1000 of the stale call; in feasibility restoration
    the restoration LP (subproblem_JuMP.jl:352-393) with the same Delta;
  * the merit function is the weighted L1 function the directional derivative of merit.jl:14 (``∇fp - μ' * cons_viol``)
    belongs to:  phi(x) = f(x) + sum_i mu_i viol_i(g(x)) + |mu|_inf * sum_j viol_j(x)   (= sum of violations in
    feasibility restoration).

Everything else is as coded: multipliers are OVERWRITTEN by the QP's each iteration (:127-128), the Hessian is evaluated
with them (sqp.jl:93), prim_infeas is the infinity norm here (:120), the second-order correction replaces a failed line
search (:196-210), the final multipliers are written back without a sign flip (:243-245).
"""
from __future__ import annotations

import numpy as np

from . import qp_solver as qs
from .sqp_tr import KT_residuals, SqpTROracle, norm_violations


class LsParameters:
    """parameters.jl:1-30, the fields the line-search driver reads."""

    def __init__(self, **kw):
        self.tol_direction = 1e-8
        self.tol_residual = 1e-8
        self.tol_infeas = 1e-8
        self.max_iter = 3000
        self.rho = 0.8
        self.eta = 0.4
        self.tau = 0.9
        self.min_alpha = 1e-6
        self.ls_delta = 1000.0  # the literal of `sub_optimize!(sqp, model, 1000.0)`, sqp_line_search.jl:255
        # fields of the trust-region Parameters the shared helpers look at
        self.init_mu = 1.0
        self.tr_size = 1000.0
        self.use_soc = False
        for k, v in kw.items():
            if not hasattr(self, k):
                raise KeyError(k)
            setattr(self, k, v)


def row_violations(E, g_L, g_U):
    return np.where(E > g_U, E - g_U, np.where(E < g_L, g_L - E, 0.0))


def norm_complementarity(E, g_L, g_U, lam):
    """common.jl:30-47 with p = Inf."""
    ineq = g_L != g_U
    with np.errstate(invalid="ignore"):
        gap = np.minimum(E - g_L, g_U - E)
    compl = np.where(ineq, gap * lam, 0.0)
    compl = np.where(np.isnan(compl), 0.0, compl)  # inf * 0 on a one-sided row whose multiplier is exactly zero
    denom = float(np.sum(lam[ineq] ** 2))
    return float(np.max(np.abs(compl), initial=0.0)) / (1.0 + np.sqrt(denom))


def weighted_merit(f, E, x, pr, mu_rows, fr):
    vr = row_violations(E, pr.g_L, pr.g_U)
    vx = float(np.sum(row_violations(x, pr.x_L, pr.x_U)))
    if fr:
        return float(np.sum(vr)) + vx
    return f + float(mu_rows @ vr) + float(np.max(np.abs(mu_rows), initial=0.0)) * vx


class SqpLSOracle(SqpTROracle):
    def __init__(self, nlp, params: LsParameters | None = None, qp_tol=1e-10):
        super().__init__(nlp, params or LsParameters(), qp_tol=qp_tol)
        pr = nlp
        # start point clamped into the bounds (:89-98; the `x_U > -Inf` test is as coded)
        self.x = np.array(pr.x0, float)
        self.x = np.where(pr.x_L > -np.inf, np.maximum(self.x, pr.x_L), self.x)
        self.x = np.where(pr.x_U > -np.inf, np.minimum(self.x, pr.x_U), self.x)
        self.soc = np.zeros(pr.n)
        self.mu_rows = np.full(pr.m, 10.0)
        self.alpha = 1.0
        self.compl = np.inf
        self.Delta = self.options.ls_delta
        self.directional_derivative = 0.0

    # --- primitives (the device computes the same quantities in k_linesearch) ---------------------------------
    def compute_mu(self):  # rule 2
        o, pr = self.options, self.problem
        if self.iter == 1:
            denom = max((1.0 - o.rho) * norm_violations(self.E, pr.g_L, pr.g_U, self.x, pr.x_L, pr.x_U, 1), 1e-8)
            hess_part = max(0.5 * float(self.p @ (self._H() @ self.p)), 0.0)
            self.mu_rows[:] = (float(self.df @ self.p) + hess_part) / denom
        else:
            self.mu_rows = np.maximum(self.mu_rows, np.abs(self.lam))

    def phi_at(self, alpha):
        pr = self.problem
        x = self.x + alpha * self.p
        f, E = self.f, self.E
        if alpha > 0.0:
            f = float(pr.eval_f(x))
            E = np.zeros(pr.m)
            pr.eval_g(x, E)
        return weighted_merit(f, E, x, pr, self.mu_rows, self.feasibility_restoration)

    def compute_derivative(self):
        pr = self.problem
        if self.feasibility_restoration:
            dfp = float(sum(np.sum(v) for v in self.p_slack.values()))
            viol = row_violations(self.E, pr.g_L, pr.g_U)
            lhs = self.E - viol
            cons = row_violations(lhs, pr.g_L, pr.g_U)
        else:
            dfp = float(self.df @ self.p)
            cons = row_violations(self.E, pr.g_L, pr.g_U)
        return dfp - float(self.mu_rows @ cons)

    def compute_alpha(self):
        o = self.options
        self.alpha = 1.0
        if np.max(np.abs(self.p), initial=0.0) <= o.tol_direction:
            return True
        phi_x_p = self.phi_at(self.alpha)
        while phi_x_p > self.phi + o.eta * self.alpha * self.directional_derivative:
            if self.alpha < o.min_alpha:
                return False
            self.alpha *= o.tau
            phi_x_p = self.phi_at(self.alpha)
        return True

    # --- run! ---------------------------------------------------------------------------------------------------
    def run(self, log=None):
        o, pr = self.options, self.problem
        self.iter = 1
        while True:
            if self.iter > o.max_iter:
                self.ret = -1
                if self.prim_infeas <= o.tol_infeas:
                    self.ret = 6
                break
            self.eval_functions()
            self.alpha = 0.0
            self.prim_infeas = norm_violations(self.E, pr.g_L, pr.g_U, self.x, pr.x_L, pr.x_U, np.inf)
            self.dual_infeas = KT_residuals(self.df, self.lam, self.mult_x_U, self.mult_x_L, self._J())
            self.compl = norm_complementarity(self.E, pr.g_L, pr.g_U, self.lam)
            self.p, self.lam, self.mult_x_U, self.mult_x_L, self.p_slack, status = self.sub_optimize()
            if status in (qs.OPTIMAL, "ALMOST_LOCALLY_SOLVED", qs.LOCALLY_SOLVED):
                pass
            elif status in (qs.INFEASIBLE, qs.LOCALLY_INFEASIBLE, "DUAL_INFEASIBLE", "NORM_LIMIT"):
                # (the reference lists INFEASIBLE only; Ipopt answers LOCALLY_INFEASIBLE -- both mean the same to the TR loop)
                if self.feasibility_restoration:
                    self.ret = 6 if self.prim_infeas <= o.tol_infeas else 2
                    break
                self.feasibility_restoration = True
                continue
            else:
                if self.prim_infeas <= o.tol_infeas:
                    self.ret = 6
                break
            self.compute_mu()
            self.phi = self.phi_at(0.0)
            self.directional_derivative = self.compute_derivative()
            is_valid_step = self.compute_alpha()
            if log is not None:
                log.append({"iter": self.iter, "fr": self.feasibility_restoration, "f": self.f, "phi": self.phi,
                            "alpha": self.alpha, "pinf": float(np.max(np.abs(self.p), initial=0.0)), "inf_pr": self.prim_infeas,
                            "inf_du": self.dual_infeas, "compl": self.compl, "mu": float(np.max(self.mu_rows, initial=0.0))})
            if np.max(np.abs(self.p), initial=0.0) <= o.tol_direction:
                if self.feasibility_restoration:
                    self.feasibility_restoration = False
                    self.iter += 1
                    continue
                self.ret = 0
                break
            if self.prim_infeas <= o.tol_infeas and self.compl <= o.tol_residual:
                if self.feasibility_restoration:
                    self.feasibility_restoration = False
                    self.iter += 1
                    continue
                elif self.dual_infeas <= o.tol_residual:
                    self.ret = 0
                    break
            if not is_valid_step:
                self.alpha = 1.0
                self.sub_optimize_soc()  # sets p_soc = p + correction (sqp_trust_region.jl:341-360)
                self.soc = self.p_soc - self.p
            self.x = self.x + self.alpha * self.p + self.soc
            self.soc = np.zeros(pr.n)
            self.iter += 1
        self.obj_val = float(pr.eval_f(self.x))
        self.status = int(self.ret)
        self.mult_g = self.lam.copy()
        return self
