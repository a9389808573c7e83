"""TEST INFRASTRUCTURE ONLY -- closed-loop replay of an SQP trajectory against the CPU oracle.

The device driver (host/sqp_trust_region.py on libsqpqp.so) runs a whole SQP solve and records every subproblem it
hands to the engine.  Every recorded subproblem is then rebuilt on the CPU from the recorded NLP values exactly as the
reference builds it (oracle/subproblem.py: subproblem_JuMP.jl:127-183, 352-393, 432-448, 465-512) and checked:

  * classification: infeasible on the device  <=>  infeasible for the oracle (HiGHS certificate standing in for Ipopt's
    failed restoration).  A mismatch is tolerated only on MARGINAL subproblems -- least l1 violation below `marginal`
    (Ipopt itself decides those by its 1e-8 tolerance) -- and counted;
  * a solved QP is a KKT point of its own data: scaled stationarity / primal / complementarity residual <= 1e-6
    (north_star's bar), in the reference's sign and storage convention (subproblem_JuMP.jl:514-563);
  * objective not worse than the oracle's solution of the same QP (convex subproblems: equal to 1e-6; nonconvex
    subproblems have several local solutions -- both are verified KKT points and the comparison is reported);
  * a solved restoration LP reaches the oracle's optimal sum of slacks to 1e-6.

This gives whole-trajectory evidence without requiring two chaotic SQP runs to stay on the same path.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
from scipy.optimize import linprog

from oracle import qp_solver as qs
from oracle.coo import CooMatrix, SymCooMatrix
from oracle.subproblem import QpData, QpOracle, trust_region_box
from sqpsolver_jl_b200 import capi

OK = (capi.MOI_OPTIMAL, 7, capi.MOI_ALMOST_LOCALLY_SOLVED, capi.MOI_LOCALLY_SOLVED)
INFEAS = (capi.MOI_INFEASIBLE, capi.MOI_LOCALLY_INFEASIBLE)


def qp_of_trace(nlp, t, b=None):
    """(P, q, A, rl, ru, xl, xu) of a recorded normal-phase subproblem; b selects per-instance bounds of a batch."""
    J = CooMatrix(nlp.j_row, nlp.j_col, nlp.m, nlp.n); J.fill(t["dE"])
    H = SymCooMatrix(nlp.h_row, nlp.h_col, nlp.n); H.fill(t["h_val"])
    gL = nlp.g_L[b] if (b is not None and np.ndim(nlp.g_L) == 2) else nlp.g_L
    gU = nlp.g_U[b] if (b is not None and np.ndim(nlp.g_U) == 2) else nlp.g_U
    lb, ub = trust_region_box(nlp.x_L - t["x"], nlp.x_U - t["x"], t["Delta"])
    return H.to_scipy(), t["df"], J.to_scipy(), gL - t["E"], gU - t["E"], lb, ub


def scaled_kkt(P, q, A, rl, ru, xl, xu, x, lam, rc):
    k = qs.kkt_residuals(P, q, A, rl, ru, xl, xu, x, lam, rc)
    sd = max(1.0, np.abs(q).max(initial=0.0), np.abs(lam).max(initial=0.0), np.abs(rc).max(initial=0.0))
    return max(k["stationarity"] / sd, k["primal"], k["complementarity"] / sd)


def least_l1_violation(A, rl, ru, xl, xu):
    """min sum of row-bound violations over the box (an LP, HiGHS): how far from feasible the constraint set is."""
    m, n = A.shape
    A = sp.csr_matrix(A)
    lo, up = np.isfinite(rl), np.isfinite(ru)
    nl, nu = int(lo.sum()), int(up.sum())
    # variables [x, sl (lo rows), su (up rows)]:  A x + sl >= rl,  A x - su <= ru
    A_ub = sp.vstack([sp.hstack([-A[lo], -sp.identity(nl), sp.csr_matrix((nl, nu))]),
                      sp.hstack([A[up], sp.csr_matrix((nu, nl)), -sp.identity(nu)])], format="csr")
    b_ub = np.concatenate([-rl[lo], ru[up]])
    c = np.concatenate([np.zeros(n), np.ones(nl + nu)])
    bounds = [(None if not np.isfinite(a) else a, None if not np.isfinite(b) else b) for a, b in zip(xl, xu)] + [(0, None)] * (nl + nu)
    res = linprog(c, A_ub=A_ub, b_ub=b_ub, bounds=bounds, method="highs")
    return float(res.fun) if res.status == 0 else np.inf


def fr_objective(nlp, t, p, b=None):
    """Sum of slacks of the restoration LP for a given p (rows satisfied at p = 0 have their slacks fixed, :365-371)."""
    J = CooMatrix(nlp.j_row, nlp.j_col, nlp.m, nlp.n); J.fill(t["dE"])
    gL = nlp.g_L[b] if (b is not None and np.ndim(nlp.g_L) == 2) else nlp.g_L
    gU = nlp.g_U[b] if (b is not None and np.ndim(nlp.g_U) == 2) else nlp.g_U
    r = J.to_scipy() @ p + t["E"]
    v = np.maximum(gL - r, 0.0) + np.maximum(r - gU, 0.0)
    return float(v[nlp.num_linear_constraints:].sum())


def check_trace(nlp, trace, oracle_every=1, marginal=1e-6, qp_tol=1e-10, verbose=False, oracle_on=None, feas_checks=None):
    """Check every recorded subproblem; the oracle solves every `oracle_every`-th one (KKT is checked on all), or exactly
    the subproblems listed in `oracle_on`.  `feas_checks` bounds the number of HiGHS feasibility verdicts on subproblems the
    device calls infeasible (a 2000-bus LP takes HiGHS tens of seconds; None = all).
    Returns a summary dict; raises AssertionError on a violated bar."""
    out = {"n": 0, "qp": 0, "fr": 0, "infeasible": 0, "marginal_mismatch": 0, "worst_kkt": 0.0, "oracle_solved": 0,
           "nonconvex": 0, "dev_better": 0, "dev_worse": 0, "same_step": 0, "almost": 0}
    for k, t in enumerate(trace):
        b = t.get("b")
        st = int(t["status"])
        out["n"] += 1
        P, q, A, rl, ru, xl, xu = qp_of_trace(nlp, t, b)
        use_oracle = (k in oracle_on) if oracle_on is not None else (k % oracle_every) == 0
        if t["fr"]:
            out["fr"] += 1
            if st in INFEAS:
                # restoration LP infeasible: only possible when the LINEAR rows / box are inconsistent
                ml = nlp.num_linear_constraints
                assert not qs.is_feasible(A[:ml], rl[:ml], ru[:ml], xl, xu), ("FR LP declared infeasible", k)
                continue
            assert st in OK, ("FR status", k, st)
            if use_oracle:
                gL = nlp.g_L[b] if (b is not None and np.ndim(nlp.g_L) == 2) else nlp.g_L
                gU = nlp.g_U[b] if (b is not None and np.ndim(nlp.g_U) == 2) else nlp.g_U
                data = QpData(P, q, A, t["E"], gL, gU, nlp.x_L, nlp.x_U, nlp.num_linear_constraints)
                ora = QpOracle(data, qp_tol=qp_tol); ora.create_model(t["Delta"])
                ro = ora.sub_optimize_FR(t["x"], t["Delta"])
                assert ro[-1] in qs.OK_STATUSES, ("oracle FR", k, ro[-1])
                o_obj, d_obj = fr_objective(nlp, t, ro[0], b), fr_objective(nlp, t, t["p"], b)
                # 1e-6 for a solve that met the full tolerance; an ALMOST_LOCALLY_SOLVED one (Ipopt's "acceptable level":
                # residuals at 1e-6, the status the reference accepts at sqp_trust_region.jl:144) is held to 1e-5
                ftol = 1e-5 if st == capi.MOI_ALMOST_LOCALLY_SOLVED else 1e-6
                assert abs(d_obj - o_obj) <= ftol * max(1.0, abs(o_obj)), ("FR optimum", k, d_obj, o_obj, st)
                out["oracle_solved"] += 1
            # the step must respect box and linear rows
            assert (t["p"] >= xl - 1e-7).all() and (t["p"] <= xu + 1e-7).all(), ("FR box", k)
            continue
        out["qp"] += 1
        if st in INFEAS:
            out["infeasible"] += 1
            assert not t["p"].any() and not t["lambda_qp"].any(), ("zero fill", k)  # collect_solution! :551-555
            if feas_checks is not None and out["infeasible"] > feas_checks:
                continue
            if qs.is_feasible(A, rl, ru, xl, xu):
                # HiGHS did not certify infeasibility (on the 2000-bus network it can also stop without a verdict): measure
                # how far from feasible the constraint set is.  Clearly positive -> infeasible, the device is right;
                # below `marginal` -> a marginal subproblem the two sides may legitimately classify differently (counted)
                v = least_l1_violation(A, rl, ru, xl, xu)
                if v <= marginal:
                    out["marginal_mismatch"] += 1
            continue
        assert st in OK, ("QP status", k, st, t["info"])
        out["almost"] += int(st == capi.MOI_ALMOST_LOCALLY_SOLVED)
        assert (t["mult_x_L"] >= 0).all() and (t["mult_x_U"] <= 0).all(), ("storage convention", k)
        kkt = scaled_kkt(P, q, A, rl, ru, xl, xu, t["p"], t["lambda_qp"], t["mult_x_L"] + t["mult_x_U"])
        out["worst_kkt"] = max(out["worst_kkt"], kkt)
        assert kkt <= 1e-6, ("scaled KKT", k, kkt, t["info"])
        if not use_oracle:
            continue
        res = qs.solve_qp(P, q, A, rl, ru, xl, xu, tol=qp_tol)
        if res.status in qs.INFEASIBLE_STATUSES:
            # the device returned a KKT point (verified above) of a QP HiGHS calls infeasible: marginal by construction
            prim = qs.kkt_residuals(P, q, A, rl, ru, xl, xu, t["p"], t["lambda_qp"], t["mult_x_L"] + t["mult_x_U"])["primal"]
            assert prim <= marginal, ("device solved, oracle infeasible", k, prim)
            out["marginal_mismatch"] += 1
            continue
        assert res.status in qs.OK_STATUSES, ("oracle status", k, res.status)
        out["oracle_solved"] += 1
        obj_d = 0.5 * t["p"] @ (P @ t["p"]) + q @ t["p"]
        tol = 1e-6 * max(1.0, abs(res.obj))
        same = np.abs(t["p"] - res.x).max() <= 1e-6 * max(1.0, np.abs(res.x).max())
        out["same_step"] += int(same)
        # convex iff the Lagrangian Hessian handed to the QP is positive semidefinite (dense check; the 2000-bus
        # network is too large for it and is treated as possibly nonconvex)
        convex = P.shape[0] <= 1500 and np.linalg.eigvalsh(P.toarray()).min() >= -1e-9 * max(1.0, abs(P).max())
        if convex:
            # one optimal value: the device must reach it
            assert obj_d <= res.obj + tol + 2.0 * np.abs(q).sum() * 1e-9, ("objective worse than the oracle's", k, obj_d, res.obj)
        else:
            out["nonconvex"] += 1
        if obj_d < res.obj - tol:
            out["dev_better"] += 1
        elif obj_d > res.obj + tol:
            out["dev_worse"] += 1
        if verbose:
            print(k, "st", st, "kkt %.1e" % kkt, "obj", obj_d, res.obj, "same" if same else "diff", "convex" if convex else "nonconvex")
    return out
