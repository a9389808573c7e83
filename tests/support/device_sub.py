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

    def create_model(self, delta):
        pr = self.driver.problem
        self.engine.setup_nlp(pr.n, pr.m, pr.num_linear_constraints, pr.j_row, pr.j_col, pr.h_row, pr.h_col, pr.x_L, pr.x_U,
                              pr.g_L, pr.g_U)
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
    driver.sub_factory = lambda data: _with_data(DeviceSub(driver, engine), data)
    return driver


def _with_data(sub, data):
    sub.data = data
    return sub
