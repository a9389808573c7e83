"""TEST INFRASTRUCTURE ONLY -- the device engine as the sub-optimizer of the ORACLE drivers.

``oracle.sqp_tr.SqpTROracle`` / ``oracle.sqp_ls.SqpLSOracle`` are line-by-line restatements of the reference's drivers
(sqp_trust_region.jl:98-223, sqp_line_search.jl:71-251) that take their QP solutions from a ``sub_factory``.  Putting
libsqpqp.so behind that hook makes a run in which the ONLY non-oracle component is the QP solve (validated per
subproblem by tests/support/closed_loop.py); comparing it with the device-side host driver (host/sqp_trust_region.py,
host/sqp_line_search.py: scatter, merit, violation norms, KT residual and line-search primitives on the device) is then
a well-posed driver-level parity test -- both runs see the same multipliers, which two different QP solvers cannot
promise on degenerate subproblems (LICQ fails on case9, the restoration LPs have non-unique minimisers).
"""
from __future__ import annotations

import numpy as np

from sqpsolver_jl_b200 import capi

NAME = {1: "OPTIMAL", 2: "INFEASIBLE", 4: "LOCALLY_SOLVED", 5: "LOCALLY_INFEASIBLE", 7: "ALMOST_OPTIMAL",
        10: "ALMOST_LOCALLY_SOLVED", 11: "ITERATION_LIMIT", 20: "NUMERICAL_ERROR"}


class DeviceSub:
    """QpJuMP's method set (subproblem_JuMP.jl:23-24, 36-183, 352-393) on the NLP lane of the C ABI, fed with the COO
    values of the oracle driver it serves (``driver.dE / h_val / df``; the row values ``data.b`` are E, or the second-
    order-correction override g(x+p) - J p of sqp_trust_region.jl:341-360)."""

    def __init__(self, driver, engine: capi.Engine):
        self.driver = driver
        self.engine = engine
        self.data = None
        self.n_solves = 0
        self.is_setup = False

    def ensure_setup(self):
        if not self.is_setup:
            pr = self.driver.problem
            self.engine.setup_nlp(pr.n, pr.m, pr.num_linear_constraints, pr.j_row, pr.j_col, pr.h_row, pr.h_col, pr.x_L, pr.x_U,
                                  pr.g_L, pr.g_U)
            self.is_setup = True

    def sub_optimize_lp(self):
        """sub_optimize_lp! (sqp_trust_region.jl:264-304) of the oracle driver with the device's start-point projection
        (phase 3) in place of the CPU oracle's: both drivers of a parity test then start from the same point."""
        d = self.driver
        pr = d.problem
        d.f = float(pr.eval_f(d.x))
        pr.eval_grad_f(d.x, d.df)
        d.eval_Jacobian()
        self.ensure_setup()
        E = np.zeros(pr.m)
        pr.eval_g(d.x, E)
        self.engine.update_nlp(d.dE, d.h_val, d.df, E)
        xs, lam, mxL, mxU, _, st, _ = self.engine.solve_tr(capi.PHASE_LP, d.x, np.inf)
        self.n_solves += 1
        d.n_qp += 1
        ok = int(st[0]) in (1, 4, 7, 10)
        lam = lam[0].copy()
        lam[pr.num_linear_constraints:] = 0.0
        outs = [xs[0].copy(), lam, mxU[0].copy(), mxL[0].copy()]
        for v in outs:
            if not ok:
                v[:] = 0.0
            v[np.abs(v) < 1e-10] = 0.0  # dropzeros! (utils.jl:16-22)
        d.x, d.lam, d.mult_x_U, d.mult_x_L = outs
        d.sub_status = NAME.get(int(st[0]), "OTHER_ERROR")

    def create_model(self, delta):
        pr = self.driver.problem
        self.ensure_setup()
        ml, m = pr.num_linear_constraints, pr.m
        self.slack_rows = []
        for i in range(ml, m):
            self.slack_rows.append(i + 1)
            if pr.g_L[i] > -np.inf and pr.g_U[i] < np.inf:
                self.slack_rows.append(i + 1)

    def _solve(self, phase, x_k, delta):
        d = self.driver
        self.engine.update_nlp(d.dE, d.h_val, d.df, self.data.b)
        p, lam, mxL, mxU, slack, st, info = self.engine.solve_tr(phase, x_k, delta)
        self.n_solves += 1
        p_slack = {}
        for c, i in enumerate(self.slack_rows):
            p_slack.setdefault(i, []).append(float(slack[0][c]))
        return p[0].copy(), lam[0].copy(), mxU[0].copy(), mxL[0].copy(), p_slack, NAME.get(int(st[0]), "OTHER_ERROR")

    def sub_optimize(self, x_k, delta):
        return self._solve(capi.PHASE_QP, x_k, delta)

    def sub_optimize_FR(self, x_k, delta):
        return self._solve(capi.PHASE_FR, x_k, delta)


def attach(driver, engine):
    """Make `driver` (an oracle SqpTROracle / SqpLSOracle) solve its subproblems on the device."""
    sub = DeviceSub(driver, engine)
    driver.sub_factory = lambda data: _with_data(sub, data)
    driver.sub_optimize_lp = sub.sub_optimize_lp  # the start-point projection goes through the device as well
    return driver


def _with_data(sub, data):
    sub.data = data
    return sub
