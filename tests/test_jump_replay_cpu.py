"""CPU check of the JuMP-replay harness itself (tests/support/jump_replay.py): with the ORACLE QP solver behind the
generic-lane calls, the replayed model of create_model! / sub_optimize! / sub_optimize_FR! / modify_constraints! /
collect_solution! (subproblem_JuMP.jl:36-183, 352-393, 465-563) must reproduce the oracle's own adapter
(oracle/subproblem.py) -- so that the GPU test that puts libsqpqp.so behind the same harness tests the engine and the
boundary, not the harness."""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "support"))

from jump_replay import JumpReplay  # noqa: E402
from oracle import qp_solver as qs  # noqa: E402
from oracle.sqp_tr import Parameters, SqpTROracle  # noqa: E402
from oracle.subproblem import QpData, QpOracle  # noqa: E402
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar  # noqa: E402
from sqpsolver_jl_b200.nlp.networks import case9  # noqa: E402
from sqpsolver_jl_b200.nlp.toy import ToyExample  # noqa: E402

CODE = {"LOCALLY_SOLVED": 4, "ALMOST_LOCALLY_SOLVED": 10, "INFEASIBLE": 2, "LOCALLY_INFEASIBLE": 5, "ITERATION_LIMIT": 11,
        "NUMERICAL_ERROR": 20}


class OracleEngine:
    """sqpqp_qp_setup / sqpqp_qp_solve semantics (include/sqpqp.h generic lane) on the CPU oracle."""

    def qp_setup(self, nv, nc, p_row, p_col, a_row, a_col):
        self.nv, self.nc = nv, nc
        self.p_row, self.p_col, self.a_row, self.a_col = (np.asarray(v) - 1 for v in (p_row, p_col, a_row, a_col))

    def qp_solve(self, p_val, q, a_val, rl, ru, cl, cu):
        nv, nc = self.nv, self.nc
        P = None
        if p_val is not None and len(p_val):
            off = self.p_row != self.p_col   # MOI triplets: off-diagonal c <=> P_ij = P_ji = c; duplicates add
            P = sp.coo_matrix((np.concatenate([p_val, p_val[off]]), (np.concatenate([self.p_row, self.p_col[off]]),
                                                                      np.concatenate([self.p_col, self.p_row[off]]))), shape=(nv, nv)).tocsr()
        A = sp.coo_matrix((a_val, (self.a_row, self.a_col)), shape=(nc, nv)).tocsr()
        res = qs.solve_qp(P, q, A, rl, ru, cl, cu)
        return res.x, res.row_dual, res.col_dual, CODE[res.status], {"objective": res.obj}


def _mixed_rows_problem():
    rng = np.random.default_rng(4)
    n, m, ml = 7, 8, 3
    A = rng.standard_normal((m, n))
    Ax = A @ rng.uniform(-0.3, 0.3, n)
    c_lb = np.array([Ax[0], Ax[1] - 0.1, -np.inf, Ax[3], Ax[4] - 0.2, -np.inf, Ax[6] + 0.0, Ax[7] - 0.05])
    c_ub = np.array([Ax[0], np.inf, Ax[2] + 0.1, Ax[3], np.inf, Ax[5] + 0.02, Ax[6] + 0.3, Ax[7] + 0.05])
    M = rng.standard_normal((n, n))
    return QpData(sp.csr_matrix(M @ M.T + np.eye(n)), 3.0 * rng.standard_normal(n), sp.csr_matrix(A), np.zeros(m), c_lb, c_ub,
                  np.full(n, -0.5), np.full(n, 0.5), ml), n, m, ml


def test_replayed_model_shape_and_dual_mapping():
    data, n, m, ml = _mixed_rows_problem()
    ora = QpOracle(data); ora.create_model(1.0)
    xo, lo, uo, Lo, _, so = ora.sub_optimize(np.zeros(n), 1.0)
    rep = JumpReplay(data, OracleEngine()); rep.create_model(1.0)
    xg, lg, ug, Lg, ps, sg = rep.sub_optimize(np.zeros(n), 1.0)
    assert so == sg == "LOCALLY_SOLVED"
    assert rep.ncol == n + (m - ml) + 3 and len(rep.constr) == m + 2 and rep.rngcons == [6, 7]
    F = rep.last["F"]
    # copy_to order: EqualTo rows first, then GreaterThan, then LessThan (incl. the appended halves)
    assert [rep.constr[k][0] for k in sorted(F["row_of"], key=F["row_of"].get)] == sorted(c[0] for c in rep.constr)
    # fixed slack columns arrive as cl = cu = 0
    assert (F["cl"][n:] == 0).all() and (F["cu"][n:] == 0).all()
    assert set(ps) == set(range(ml + 1, m + 1))
    assert np.abs(xg - xo).max() <= 1e-7
    assert np.abs(lg - lo).max() <= 1e-6 * max(1.0, np.abs(lo).max())
    assert np.abs((ug + Lg) - (uo + Lo)).max() <= 1e-6 * max(1.0, np.abs(uo + Lo).max())
    # restoration phase: rows violated at p = 0 get free slacks, the optimum equals the oracle adapter's
    data.b = data.b + 0.4 * np.sign(np.arange(m) - 3.5)
    ro = ora.sub_optimize_FR(np.zeros(n), 1.0)
    rg = rep.sub_optimize_FR(np.zeros(n), 1.0)
    so_, sg_ = sum(sum(v) for v in ro[4].values()), sum(sum(v) for v in rg[4].values())
    assert ro[-1] in qs.OK_STATUSES and rg[-1] in qs.OK_STATUSES and abs(so_ - sg_) <= 1e-7 * max(1.0, so_)


def test_sqp_through_the_replayed_model_equals_the_oracle_adapter():
    for make, kw in ((ToyExample, dict(max_iter=100)), (lambda: AcopfPolar(case9()), dict(max_iter=100, init_mu=1e4))):
        a = SqpTROracle(make(), Parameters(**kw)).run()
        b = SqpTROracle(make(), Parameters(**kw), sub_factory=lambda data: JumpReplay(data, OracleEngine())).run()
        assert a.status == b.status == 0 and a.iter == b.iter
        assert abs(a.obj_val - b.obj_val) <= 1e-8 * max(1.0, abs(a.obj_val))
