"""Host mirror of the reference's SQP line-search driver on the device engine.

``SqpLS.run`` restates ``run!(::SqpLS)`` (sqp_line_search.jl:71-251) with the QP subproblem solved by the CUDA engine
(``QpDevice.sub_optimize`` / ``sub_optimize_FR`` / ``sub_optimize_soc``) and the merit arithmetic -- penalty rule 2
(:280-291), Armijo step (:303-334), merit value (sqp.jl:170-183), directional derivative (sqp.jl:190-213,
merit.jl:13-17), complementarity (common.jl:30-47), violation norms (common.jl:54-77), KT residual (common.jl:14-23)
-- by ``sqpqp_linesearch_terms`` / ``sqpqp_kt_residuals`` on the device.

The reference does not compile this driver (``# include("sqp_line_search.jl")``, sqp.jl:226) and the file is stale; the
two gaps (the non-existent three-argument ``sub_optimize!(sqp, model, 1000.0)`` and the vector-valued ``compute_phi``) are
closed exactly as in the CPU restatement ``oracle/sqp_ls.py`` -- see its header; the parity tests compare the two.
"""
from __future__ import annotations

import numpy as np

from .subproblem import INFEASIBLE_STATUSES, OK_STATUSES, QpDevice


class LsParameters:
    """parameters.jl:1-30, the fields the line-search driver reads (defaults of the reference)."""

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
        for k, v in kw.items():
            if not hasattr(self, k):
                raise KeyError(k)
            setattr(self, k, v)


class SqpLS:
    def __init__(self, nlp, params: LsParameters | None = None, device: int = 0, engine_options: dict | None = None):
        self.problem = nlp
        self.options = params or LsParameters()
        n, m = nlp.n, nlp.m
        x = np.array(nlp.x0, float)
        x = np.where(nlp.x_L > -np.inf, np.maximum(x, nlp.x_L), x)  # :89-98 (`x_U > -Inf` as coded)
        x = np.where(nlp.x_U > -np.inf, np.minimum(x, nlp.x_U), x)
        self.x = x
        self.p = np.zeros(n)
        self.soc = np.zeros(n)
        self.p_slack = np.zeros(0)
        self.lam = np.zeros(m)
        self.mult_x_L = np.zeros(n)
        self.mult_x_U = np.zeros(n)
        self.df = np.zeros((1, n))
        self.E = np.zeros((1, m))
        self.dE = np.zeros((1, nlp.nnz_jac_coo))
        self.h_val = np.zeros((1, nlp.nnz_hess_coo))
        self.mu_rows = np.full(m, 10.0)
        self.phi = np.inf
        self.alpha = 1.0
        self.prim_infeas = self.dual_infeas = self.compl = np.inf
        self.feasibility_restoration = False
        self.iter = 1
        self.ret = -5
        self.n_qp = 0
        self.optimizer = QpDevice(nlp, batch=1, device=device, engine_options=engine_options)
        self.optimizer.create_model()
        self.engine = self.optimizer.engine

    # ---- eval_functions! (sqp.jl:86-117): host callbacks, then the device scatter ----------------------------
    def eval_functions(self):
        pr = self.problem
        X = self.x[None, :]
        self.f = float(np.asarray(pr.eval_f(X)).reshape(-1)[0])
        pr.eval_grad_f(X, self.df)
        pr.eval_g(X, self.E)
        pr.eval_jac_g(X, self.dE)
        pr.eval_h(X, 1.0, self.lam[None, :], self.h_val)
        self.optimizer.update(self.dE, self.h_val, self.df, self.E)

    def _terms(self, alpha):
        """Device primitives at x + alpha p (f, g at the trial point come from the host callbacks, sqp.jl:175-178)."""
        pr = self.problem
        if alpha > 0.0:
            xt = (self.x + alpha * self.p)[None, :]
            f = float(np.asarray(pr.eval_f(xt)).reshape(-1)[0])
            Et = np.zeros((1, pr.m))
            pr.eval_g(xt, Et)
        else:
            f, Et = self.f, self.E
        t = self.engine.linesearch_terms(self.x, self.p, alpha, Et, self.mu_rows, self.lam)
        return f, {k: float(v[0]) for k, v in t.items()}

    def phi_at(self, alpha):
        f, t = self._terms(alpha)
        if self.feasibility_restoration:
            return t["viol1_trial"]
        return f + t["wviol_trial"]

    def compute_mu(self, t0):  # rule 2, sqp_line_search.jl:280-291
        o = self.options
        if self.iter == 1:
            denom = max((1.0 - o.rho) * t0["viol1"], 1e-8)
            hess_part = max(0.5 * t0["pHp"], 0.0)
            self.mu_rows[:] = (t0["dfp"] + hess_part) / denom
        else:
            self.mu_rows = np.maximum(self.mu_rows, np.abs(self.lam))

    def compute_derivative(self, t0):  # sqp.jl:190-213 + merit.jl:14
        if self.feasibility_restoration:
            # cons_viol of the restoration branch is max(0, lhs - g_U, g_L - lhs) with lhs = E - viol: zero on
            # satisfied rows and viol on rows violated from above only -> evaluated on the host from E (m values)
            pr = self.problem
            E = self.E[0]
            viol = np.where(E > pr.g_U, E - pr.g_U, np.where(E < pr.g_L, pr.g_L - E, 0.0))
            lhs = E - viol
            cons = np.where(lhs > pr.g_U, lhs - pr.g_U, np.where(lhs < pr.g_L, pr.g_L - lhs, 0.0))
            return float(np.sum(self.p_slack)) - float(self.mu_rows @ cons)
        # weighted row violation at x: wviol0 minus the bound part (zero: x stays inside its bounds in this driver)
        pr = self.problem
        vx = float(np.sum(np.where(self.x > pr.x_U, self.x - pr.x_U, np.where(self.x < pr.x_L, pr.x_L - self.x, 0.0))))
        return t0["dfp"] - (t0["wviol0"] - float(np.max(np.abs(self.mu_rows), initial=0.0)) * vx)

    def compute_alpha(self):  # :303-334
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

    def run(self, log=None):
        o, pr = self.options, self.problem
        delta = np.full(1, o.ls_delta)
        self.iter = 1
        while True:
            if self.iter > o.max_iter:
                self.ret = -1
                if self.prim_infeas <= o.tol_infeas:
                    self.ret = 6
                break
            self.eval_functions()
            self.alpha = 0.0
            self.p[:] = 0.0
            _, t = self._terms(0.0)  # at the current point with the multipliers of the previous QP
            self.prim_infeas = t["violinf"]
            self.compl = t["compl"]
            self.dual_infeas = float(self.engine.kt_residuals(self.lam, self.mult_x_U, self.mult_x_L)[0])
            X = self.x[None, :]
            if self.feasibility_restoration:
                p, lam, mxU, mxL, slack, status = self.optimizer.sub_optimize_FR(X, delta)
            else:
                p, lam, mxU, mxL, slack, status = self.optimizer.sub_optimize(X, delta)
            self.n_qp += 1
            st = int(status[0])
            self.p, self.lam, self.mult_x_U, self.mult_x_L = p[0].copy(), lam[0].copy(), mxU[0].copy(), mxL[0].copy()
            self.p_slack = slack[0].copy()
            if st in OK_STATUSES:
                pass
            elif st in INFEASIBLE_STATUSES:
                if self.feasibility_restoration:
                    self.ret = 6 if self.prim_infeas <= o.tol_infeas else 2
                    break
                self.feasibility_restoration = True
                continue
            else:
                if self.prim_infeas <= o.tol_infeas:
                    self.ret = 6
                break
            _, t0 = self._terms(0.0)
            self.compute_mu(t0)
            _, t0 = self._terms(0.0)  # weighted violation with the new penalties
            self.phi = t0["viol1_trial"] if self.feasibility_restoration else self.f + t0["wviol_trial"]
            self.directional_derivative = self.compute_derivative(t0)
            is_valid_step = self.compute_alpha()
            pinf = float(np.max(np.abs(self.p), initial=0.0))
            if log is not None:
                log.append({"iter": self.iter, "fr": self.feasibility_restoration, "f": self.f, "phi": self.phi,
                            "alpha": self.alpha, "pinf": pinf, "inf_pr": self.prim_infeas, "inf_du": self.dual_infeas,
                            "compl": self.compl, "mu": float(np.max(self.mu_rows, initial=0.0)), "sub_status": st})
            if pinf <= o.tol_direction:
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
            if not is_valid_step:  # second-order correction replaces the failed line search (:196-210)
                self.alpha = 1.0
                E_soc = np.zeros((1, pr.m))
                pr.eval_g((self.x + self.p)[None, :], E_soc)
                E_soc -= self.engine.jac_times(self.p)
                ps, *_ = self.optimizer.sub_optimize_soc(X, delta, E_soc)
                self.n_qp += 1
                self.soc = ps[0].copy()
            self.x = self.x + self.alpha * self.p + self.soc
            self.soc = np.zeros(pr.n)
            self.iter += 1
        self.obj_val = float(np.asarray(pr.eval_f(self.x[None, :])).reshape(-1)[0])
        self.status = int(self.ret)
        self.mult_g = self.lam.copy()
        return self

    def close(self):
        self.optimizer.close()
