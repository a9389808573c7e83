"""TEST INFRASTRUCTURE ONLY -- replay of the JuMP model the reference's QP adapter really builds, driven through the
GENERIC lane of the C ABI (boundary B1: what an ``MOI.AbstractOptimizer`` shim sees).

Julia/JuMP cannot run in the build image, so the traffic is restated here from the reference's code:

  create_model!        subproblem_JuMP.jl:36-125   n + S columns (x, then per nonlinear row u_i and -- if both bounds are
                                                   finite -- v_i, lower bound 0); one row per constraint, the `<=` half
                                                   of a two-sided row appended at index m + k (:116-124); nonlinear rows
                                                   start with their slack terms only (:87,95,103,110)
  sub_optimize!        :127-183                    objective rebuilt (sum c_i x_i + 1/2 sum Q_ij x_i x_j), every slack
                                                   fixed to 0, trust-region box, coefficients + right-hand sides
  sub_optimize_FR!     :352-393                    objective = sum of slacks; slacks of rows with c_lb <= b <= c_ub
                                                   fixed to 0, the others freed with lower bound 0
  set_trust_region!    :432-448
  modify_constraints!  :465-512                    per-nonzero coefficients of rows > m_lin, the paired row again (:482-489)
  collect_solution!    :514-563                    lambda[i] = dual(constr[i]) (+ dual(constr[m+k]) for two-sided rows),
                                                   reduced cost split into mult_x_L (> 0) / mult_x_U (< 0), p_slack as
                                                   Dict row -> values

``flatten`` is the Python twin of ``SqpQpB200.copy_to`` (sqpsolver.jl_b200/julia/SqpQpB200.jl): variables in creation
order; affine rows grouped by set type in the order EqualTo, GreaterThan, LessThan (the order JuMP's
``ListOfConstraintIndices`` loop visits them), variable bounds folded into cl / cu; quadratic objective as MOI
``ScalarQuadraticTerm`` triplets (one triangle: JuMP merges the (i,j)/(j,i) halves of the reference's full-symmetric sum;
an off-diagonal coefficient c means P_ij = P_ji = c, a diagonal one P_ii = c).  The duals are read back through the same
index map the shim keeps (``dest.rows``).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from sqpsolver_jl_b200 import capi

INF = np.inf
EQ, GE, LE = 0, 1, 2  # MOI.EqualTo, MOI.GreaterThan, MOI.LessThan
OK = (capi.MOI_OPTIMAL, 7, capi.MOI_ALMOST_LOCALLY_SOLVED, capi.MOI_LOCALLY_SOLVED)
INFEAS = (capi.MOI_INFEASIBLE, capi.MOI_LOCALLY_INFEASIBLE)
STATUS_NAME = {1: "OPTIMAL", 2: "INFEASIBLE", 4: "LOCALLY_SOLVED", 5: "LOCALLY_INFEASIBLE", 7: "ALMOST_OPTIMAL",
               10: "ALMOST_LOCALLY_SOLVED", 11: "ITERATION_LIMIT", 20: "NUMERICAL_ERROR"}


class JumpReplay:
    """Same method set and return tuples as ``QpJuMP`` / ``oracle.subproblem.QpOracle``; usable as the ``sub_factory``
    of ``oracle.sqp_tr.SqpTROracle`` so that a whole SQP solve runs through the generic lane."""

    def __init__(self, data, engine: capi.Engine):
        self.data = data
        self.engine = engine
        self.pattern_key = None
        self.n_setups = 0
        self.n_solves = 0

    # ------------------------------------------------------------------ create_model!
    def create_model(self, delta):
        d = self.data
        n, m, ml = d.c.shape[0], d.c_lb.shape[0], d.num_linear_constraints
        self.n, self.m = n, m
        A = sp.csr_matrix(d.A)
        self.col_lb = np.full(n, -INF)
        self.col_ub = np.full(n, INF)
        self.col_fixed = np.zeros(n, bool)
        self.slack_vars = {}  # row (0-based) -> list of column indices
        for i in range(ml, m):
            self.slack_vars[i] = [self._add_col(0.0)]
            if d.c_lb[i] > -INF and d.c_ub[i] < INF:
                self.slack_vars[i].append(self._add_col(0.0))
        self.constr = []   # list of [set type, {col: coef}, rhs]
        self.rngcons = []
        for i in range(m):
            lin = i < ml
            xrow = {int(j): float(v) for j, v in zip(A.indices[A.indptr[i]:A.indptr[i + 1]], A.data[A.indptr[i]:A.indptr[i + 1]])}
            lo, up = d.c_lb[i] > -INF, d.c_ub[i] < INF
            if d.c_lb[i] == d.c_ub[i]:
                self.constr.append([EQ, dict(xrow) if lin else {self.slack_vars[i][0]: 1.0, self.slack_vars[i][1]: -1.0}, 0.0])
            elif lo and up:
                self.constr.append([GE, dict(xrow) if lin else {self.slack_vars[i][0]: 1.0}, 0.0])
                self.rngcons.append(i)
            elif lo:
                self.constr.append([GE, dict(xrow) if lin else {self.slack_vars[i][0]: 1.0}, 0.0])
            elif up:
                self.constr.append([LE, dict(xrow) if lin else {self.slack_vars[i][0]: -1.0}, 0.0])
            else:
                raise ValueError("free row: the reference creates no constraint for it and its constr[] indexing breaks")
        for i in self.rngcons:
            lin = i < ml
            xrow = {int(j): float(v) for j, v in zip(A.indices[A.indptr[i]:A.indptr[i + 1]], A.data[A.indptr[i]:A.indptr[i + 1]])}
            self.constr.append([LE, dict(xrow) if lin else {self.slack_vars[i][1]: -1.0}, 0.0])
        self.obj_quad = {}
        self.obj_lin = np.zeros(self.ncol)

    def _add_col(self, lb):
        self.col_lb = np.append(self.col_lb, lb)
        self.col_ub = np.append(self.col_ub, INF)
        self.col_fixed = np.append(self.col_fixed, False)
        return self.col_lb.shape[0] - 1

    @property
    def ncol(self):
        return self.col_lb.shape[0]

    # ------------------------------------------------------------------ model edits
    def _set_trust_region(self, x_k, delta):
        d = self.data
        v_lb, v_ub = d.v_lb - x_k, d.v_ub - x_k
        for i in range(self.n):
            lb, ub = max(-delta, v_lb[i]), min(delta, v_ub[i])
            if lb > ub:
                lb, ub = max(-delta, min(0.0, v_lb[i])), min(delta, max(0.0, v_ub[i]))
            self.col_lb[i], self.col_ub[i] = lb, ub

    def _modify_constraints(self):
        d = self.data
        m, ml = self.m, d.num_linear_constraints
        A = sp.csc_matrix(d.A)
        for j in range(A.shape[1]):
            for k in range(A.indptr[j], A.indptr[j + 1]):
                i = int(A.indices[k])
                if i >= ml:
                    self.constr[i][1][j] = float(A.data[k])
        Ar = sp.csr_matrix(d.A)
        for ind, i in enumerate(self.rngcons):
            if i >= ml:
                for j, v in zip(Ar.indices[Ar.indptr[i]:Ar.indptr[i + 1]], Ar.data[Ar.indptr[i]:Ar.indptr[i + 1]]):
                    self.constr[m + ind][1][int(j)] = float(v)
        for i in range(m):
            c_ub, c_lb = d.c_ub[i] - d.b[i], d.c_lb[i] - d.b[i]
            if d.c_lb[i] == d.c_ub[i] or d.c_lb[i] > -INF:
                self.constr[i][2] = c_lb
            elif d.c_ub[i] < INF:
                self.constr[i][2] = c_ub
        for ind, i in enumerate(self.rngcons):
            self.constr[m + ind][2] = d.c_ub[i] - d.b[i]

    # ------------------------------------------------------------------ copy_to + optimize!
    def flatten(self):
        nv = self.ncol
        order = [k for st in (EQ, GE, LE) for k, c in enumerate(self.constr) if c[0] == st]
        row_of = {k: r for r, k in enumerate(order)}
        a_row, a_col, a_val, rl, ru = [], [], [], [], []
        for r, k in enumerate(order):
            st, coef, rhs = self.constr[k]
            rl.append(rhs if st in (EQ, GE) else -INF)
            ru.append(rhs if st in (EQ, LE) else INF)
            for j in sorted(coef):
                a_row.append(r + 1); a_col.append(j + 1); a_val.append(coef[j])
        cl = np.where(self.col_fixed, 0.0, self.col_lb)
        cu = np.where(self.col_fixed, 0.0, self.col_ub)
        keys = sorted(self.obj_quad)
        p_row = [i + 1 for i, _ in keys]
        p_col = [j + 1 for _, j in keys]
        p_val = [self.obj_quad[k] for k in keys]
        return dict(nv=nv, nc=len(order), p_row=np.array(p_row, np.int64), p_col=np.array(p_col, np.int64),
                    p_val=np.array(p_val, float), q=self.obj_lin.copy(), a_row=np.array(a_row, np.int64),
                    a_col=np.array(a_col, np.int64), a_val=np.array(a_val, float), rl=np.array(rl, float),
                    ru=np.array(ru, float), cl=cl, cu=cu, row_of=row_of)

    def _optimize(self):
        F = self.flatten()
        key = (F["nv"], F["nc"], F["p_row"].tobytes(), F["p_col"].tobytes(), F["a_row"].tobytes(), F["a_col"].tobytes())
        if key != self.pattern_key:  # same pattern across SQP iterations -> the device structure is kept
            self.engine.qp_setup(F["nv"], F["nc"], F["p_row"], F["p_col"], F["a_row"], F["a_col"])
            self.pattern_key = key
            self.n_setups += 1
        x, rd, cd, st, info = self.engine.qp_solve(F["p_val"] if F["p_val"].size else None, F["q"], F["a_val"], F["rl"], F["ru"],
                                                   F["cl"], F["cu"])
        self.n_solves += 1
        self.last = dict(F=F, x=x, row_dual=rd, col_dual=cd, status=st, info=info)
        return F, x, rd, cd, st

    # ------------------------------------------------------------------ sub_optimize! / sub_optimize_FR!
    def sub_optimize(self, x_k, delta):
        d = self.data
        n = self.n
        self.obj_lin = np.zeros(self.ncol)
        self.obj_lin[:n] = d.c
        self.obj_quad = {}
        if d.Q is not None:
            Q = sp.coo_matrix(d.Q)
            for i, j, v in zip(Q.row, Q.col, Q.data):  # 1/2 sum Q_ij x_i x_j over the full-symmetric nzval; JuMP merges halves
                a, b = (int(i), int(j)) if i <= j else (int(j), int(i))
                self.obj_quad[(a, b)] = self.obj_quad.get((a, b), 0.0) + (float(v) if i == j else 0.5 * float(v))
        for slacks in self.slack_vars.values():
            for s in slacks:
                self.col_lb[s] = -INF  # delete_lower_bound; fix(s, 0.0)
                self.col_fixed[s] = True
        self._set_trust_region(x_k, delta)
        self._modify_constraints()
        return self._collect(*self._optimize())

    def sub_optimize_FR(self, x_k, delta):
        d = self.data
        self.obj_quad = {}
        self.obj_lin = np.zeros(self.ncol)
        for i, slacks in self.slack_vars.items():
            feas = d.c_lb[i] <= d.b[i] <= d.c_ub[i]
            for s in slacks:
                self.obj_lin[s] = 1.0
                if feas:
                    self.col_fixed[s] = True
                else:
                    self.col_fixed[s] = False
                    self.col_lb[s] = 0.0
        self._set_trust_region(x_k, delta)
        self._modify_constraints()
        return self._collect(*self._optimize())

    # ------------------------------------------------------------------ collect_solution!
    def _collect(self, F, x, rd, cd, st):
        n, m = self.n, self.m
        mult_x_U, mult_x_L = np.zeros(n), np.zeros(n)
        p_slack = {}
        status = STATUS_NAME.get(int(st), "OTHER_ERROR")
        if st in OK:
            Xsol = x[:n].copy()
            for i, slacks in self.slack_vars.items():
                p_slack[i + 1] = [float(x[s]) for s in slacks]
            lam = np.array([rd[F["row_of"][i]] for i in range(m)])
            for ind, i in enumerate(self.rngcons):
                lam[i] += rd[F["row_of"][m + ind]]
            rc = cd[:n]  # lower- and upper-bound constraint duals of a variable summed (JuMP.reduced_cost)
            mult_x_L = np.where(rc > 0, rc, 0.0)
            mult_x_U = np.where(rc < 0, rc, 0.0)
        elif st in INFEAS:
            Xsol, lam = np.zeros(n), np.zeros(m)
        else:
            Xsol, lam = np.full(n, np.nan), np.full(m, np.nan)
        return Xsol, lam, mult_x_U, mult_x_L, p_slack, status
