"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's QP-subproblem adapter.

Follows src/algorithms/subproblem.jl:3-23 (the QP form / ``QpData``) and
src/algorithms/subproblem_JuMP.jl:

  create_model!        :36-125   rows, slack columns, range-row pairing
  sub_optimize!        :127-183  normal phase (all slacks fixed to 0)
  sub_optimize_FR!     :352-393  feasibility restoration LP (min sum of slacks)
  sub_optimize_lp      :185-244  start-point projection  min sum (x - x_k)^2
  set_trust_region!    :432-463  box rule incl. the lb > ub fallback
  modify_constraints!  :465-512  row coefficients and right-hand sides
  collect_solution!    :514-563  primal/dual read-back, reduced-cost split

The JuMP model has n + S columns and m + R rows (R = #range rows, appended at
index m+k).  Mathematically a paired (>=, <=) couple over the same expression is
one two-sided row, and ``lambda[i] = dual(>= part) + dual(<= part)``
(:537-539) is that row's single multiplier, so the restatement hands the solver
two-sided rows.  In the FR phase a range row has *different* slack columns in
its two halves (``J p + u >= lo`` and ``J p - v <= hi``, :95,122), which is kept
exactly: the two halves stay separate rows there.

Sub-solver: :mod:`oracle.qp_solver` (our stand-in for Ipopt -- parity unpinned).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from . import qp_solver as qs

INF = np.inf


class QpData:
    """subproblem.jl:12-23.  Q/A are scipy CSR; ``b`` is the constraint value E."""

    def __init__(self, Q, c, A, b, c_lb, c_ub, v_lb, v_ub, num_linear_constraints):
        self.Q, self.c, self.A, self.b = Q, c, A, b
        self.c_lb, self.c_ub, self.v_lb, self.v_ub = c_lb, c_ub, v_lb, v_ub
        self.num_linear_constraints = num_linear_constraints


def trust_region_box(v_lb, v_ub, delta):
    """set_trust_region! (subproblem_JuMP.jl:432-448), vectorised."""
    lb = np.maximum(-delta, v_lb)
    ub = np.minimum(+delta, v_ub)
    bad = lb > ub
    if bad.any():
        lb = np.where(bad, np.maximum(-delta, np.minimum(0.0, v_lb)), lb)
        ub = np.where(bad, np.minimum(+delta, np.maximum(0.0, v_ub)), ub)
    return lb, ub


def row_kinds(c_lb, c_ub):
    """Row classification used throughout create_model!/modify_constraints! (:79-112)."""
    eq = c_lb == c_ub
    rng = ~eq & (c_lb > -INF) & (c_ub < INF)
    lo = ~eq & ~rng & (c_lb > -INF)
    up = ~eq & ~rng & ~lo & (c_ub < INF)
    return eq, rng, lo, up


class QpOracle:
    """Mirror of ``QpJuMP <: AbstractSubOptimizer`` driving :func:`qp_solver.solve_qp`."""

    def __init__(self, data: QpData, qp_tol=1e-10):
        self.data = data
        self.qp_tol = qp_tol
        self.n_solves = 0
        self.last = None

    # create_model! only fixes the column/row layout; nothing to persist here.
    def create_model(self, delta):
        d = self.data
        m = d.c_lb.shape[0]
        self.slack_rows = np.arange(d.num_linear_constraints, m)
        eq, rng, lo, up = row_kinds(d.c_lb, d.c_ub)
        # 2 slack columns iff both bounds finite (:62-64) -- equality or range rows
        self.two_slacks = (d.c_lb > -INF) & (d.c_ub < INF)

    # ---------------------------------------------------------------- normal phase
    def sub_optimize(self, x_k, delta):
        d = self.data
        n = d.c.shape[0]
        m = d.c_lb.shape[0]
        lb, ub = trust_region_box(d.v_lb - x_k, d.v_ub - x_k, delta)
        rl = d.c_lb - d.b
        ru = d.c_ub - d.b
        res = qs.solve_qp(d.Q, d.c, d.A, rl, ru, lb, ub, tol=self.qp_tol)
        self.n_solves += 1
        self.last = res
        return self._collect(res, n, m, slack_values=None)

    # -------------------------------------------------------- feasibility restoration
    def sub_optimize_FR(self, x_k, delta):
        d = self.data
        n = d.c.shape[0]
        m = d.c_lb.shape[0]
        mlin = d.num_linear_constraints
        lb, ub = trust_region_box(d.v_lb - x_k, d.v_ub - x_k, delta)
        eq, rng, lo, up = row_kinds(d.c_lb, d.c_ub)
        feas = (d.b >= d.c_lb) & (d.b <= d.c_ub)  # :366 -> slacks of these rows fixed to 0
        A = sp.csr_matrix(d.A)
        rows_A, rl, ru = [], [], []
        s_cols, s_rows, s_sign = [], [], []  # slack column -> (solver row, +-1)
        owner = []  # (reference row i, which slack 0/1)
        r = 0
        row_map = []  # solver row -> reference row (for lambda accumulation)
        for i in range(m):
            clb, cub = d.c_lb[i] - d.b[i], d.c_ub[i] - d.b[i]
            nonlin = i >= mlin
            if eq[i]:
                rows_A.append(i); rl.append(clb); ru.append(clb); row_map.append(i)
                if nonlin:
                    s_rows += [r, r]; s_sign += [1.0, -1.0]; owner += [(i, 0), (i, 1)]
                r += 1
            elif rng[i]:
                rows_A.append(i); rl.append(clb); ru.append(INF); row_map.append(i)
                if nonlin:
                    s_rows.append(r); s_sign.append(1.0); owner.append((i, 0))
                r += 1
                rows_A.append(i); rl.append(-INF); ru.append(cub); row_map.append(i)
                if nonlin:
                    s_rows.append(r); s_sign.append(-1.0); owner.append((i, 1))
                r += 1
            elif lo[i]:
                rows_A.append(i); rl.append(clb); ru.append(INF); row_map.append(i)
                if nonlin:
                    s_rows.append(r); s_sign.append(1.0); owner.append((i, 0))
                r += 1
            elif up[i]:
                rows_A.append(i); rl.append(-INF); ru.append(cub); row_map.append(i)
                if nonlin:
                    s_rows.append(r); s_sign.append(-1.0); owner.append((i, 0))
                r += 1
            else:  # free row: no constraint is created for it in the reference either
                pass
        S = len(s_rows)
        Arows = A[rows_A]
        Smat = sp.csr_matrix((s_sign, (s_rows, np.arange(S))), shape=(r, S))
        Afull = sp.hstack([Arows, Smat], format="csr")
        s_free = np.array([not feas[i] for (i, _) in owner], dtype=bool)
        xl = np.concatenate([lb, np.zeros(S)])
        xu = np.concatenate([ub, np.where(s_free, INF, 0.0)])
        q = np.concatenate([np.zeros(n), np.ones(S)])
        res = qs.solve_qp(None, q, Afull, np.array(rl), np.array(ru), xl, xu, tol=self.qp_tol)
        self.n_solves += 1
        self.last = res
        # fold the solver rows back onto the reference rows
        lam = np.zeros(m)
        np.add.at(lam, np.array(row_map, dtype=int), res.row_dual)
        folded = qs.QpResult(res.status, res.x[:n], lam, res.col_dual[:n], res.obj, res.iters, res.kkt_error)
        p_slack = {}
        for c, (i, w) in enumerate(owner):
            p_slack.setdefault(i + 1, []).append(res.x[n + c])
        return self._collect(folded, n, m, slack_values=p_slack)

    # ------------------------------------------------------------------ collect
    def _collect(self, res, n, m, slack_values):
        """collect_solution! (:514-563)."""
        status = res.status
        mult_x_U = np.zeros(n)
        mult_x_L = np.zeros(n)
        p_slack = {}
        if status in qs.OK_STATUSES:
            Xsol = res.x[:n].copy()
            lam = res.row_dual.copy()
            rc = res.col_dual[:n]
            mult_x_L = np.where(rc > 0, rc, 0.0)
            mult_x_U = np.where(rc < 0, rc, 0.0)
            if slack_values is None:
                d = self.data
                for i in range(d.num_linear_constraints, m):
                    p_slack[i + 1] = [0.0, 0.0] if self.two_slacks[i] else [0.0]
            else:
                p_slack = slack_values
        elif status in qs.INFEASIBLE_STATUSES + ("DUAL_INFEASIBLE", "NORM_LIMIT", "OBJECTIVE_LIMIT"):
            Xsol = np.zeros(n)
            lam = np.zeros(m)
        else:  # ITERATION_LIMIT / unexpected: the reference returns uninitialised memory
            Xsol = np.full(n, np.nan)
            lam = np.full(m, np.nan)
        return Xsol, lam, mult_x_U, mult_x_L, p_slack, status


def sub_optimize_lp(A, cl, cu, xl, xu, x_k, m_lin, num_constraints, qp_tol=1e-10):
    """Start-point projection, subproblem_JuMP.jl:185-244.

    ``min sum (x_i - x_k_i)^2`` over the first ``m_lin`` rows of A and the
    variable bounds; returns x, lambda (length num_constraints), mult_x_U,
    mult_x_L, status.
    """
    n = x_k.shape[0]
    A = sp.csr_matrix(A)[:m_lin]
    P = 2.0 * sp.identity(n, format="csr")
    q = -2.0 * x_k
    res = qs.solve_qp(P, q, A, cl[:m_lin], cu[:m_lin], xl, xu, x0=x_k, tol=qp_tol)
    lam = np.zeros(num_constraints)
    mult_x_U = np.zeros(n)
    mult_x_L = np.zeros(n)
    if res.status in qs.OK_STATUSES:
        Xsol = res.x.copy()
        lam[:m_lin] = res.row_dual
        mult_x_L = np.where(res.col_dual > 0, res.col_dual, 0.0)
        mult_x_U = np.where(res.col_dual < 0, res.col_dual, 0.0)
    else:
        Xsol = np.zeros(n)
    return Xsol, lam, mult_x_U, mult_x_L, res.status
