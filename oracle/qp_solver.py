"""TEST INFRASTRUCTURE ONLY -- the CPU stand-in for the reference's external QP solver.

PARITY UNPINNED at the QP level: the reference delegates every QP/LP subproblem
to Ipopt (test/ext_solver.jl:2-6, examples/toy_example.jl:4-8, test/opf.jl:13-17),
an un-vendored, un-versioned dependency (no Manifest.toml, .gitignore:3) that is
not installed here, and no reference test inspects a subproblem quantity
(test/MOI_wrapper.jl:49-54 even excludes ConstraintDual).  This file therefore
restates the *published algorithm* Ipopt implements -- a primal-dual
log-barrier interior-point method with fraction-to-the-boundary steps, a
monotone (Fiacco-McCormick) barrier update, inertia/curvature correction of the
augmented system and an l1-merit backtracking line search (Waechter & Biegler,
Math. Prog. 106, 2006, sections 2-3) -- specialised to

    min 1/2 x'Px + q'x   s.t.  rl <= A x <= ru,  xl <= x <= xu

and returns the solution in the MOI conventions the reference reads back in
subproblem_JuMP.jl:514-563: row duals with  grad = A' lambda + r  (>= rows
lambda >= 0, <= rows lambda <= 0) and reduced costs r (r > 0 at a lower bound).
Primal infeasibility is certified independently with HiGHS (scipy.linprog)
on the constraint set, standing in for Ipopt's restoration-phase failure
(LOCALLY_INFEASIBLE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.optimize import linprog

# MOI.TerminationStatusCode values the SQP driver branches on
OPTIMAL = "OPTIMAL"
LOCALLY_SOLVED = "LOCALLY_SOLVED"
INFEASIBLE = "INFEASIBLE"
LOCALLY_INFEASIBLE = "LOCALLY_INFEASIBLE"
ITERATION_LIMIT = "ITERATION_LIMIT"
NUMERICAL_ERROR = "NUMERICAL_ERROR"

OK_STATUSES = (OPTIMAL, "ALMOST_OPTIMAL", "ALMOST_LOCALLY_SOLVED", LOCALLY_SOLVED)
INFEASIBLE_STATUSES = (INFEASIBLE, LOCALLY_INFEASIBLE)


class QpResult:
    __slots__ = ("status", "x", "row_dual", "col_dual", "obj", "iters", "kkt_error")

    def __init__(self, status, x, row_dual, col_dual, obj=np.nan, iters=0, kkt_error=np.nan):
        self.status, self.x, self.row_dual, self.col_dual = status, x, row_dual, col_dual
        self.obj, self.iters, self.kkt_error = obj, iters, kkt_error


def is_feasible(A, rl, ru, xl, xu, tol=1e-9):
    """HiGHS feasibility check of {rl <= Ax <= ru, xl <= x <= xu}."""
    m, n = A.shape
    A = sp.csr_matrix(A)
    eq = np.isfinite(rl) & (rl == ru)
    up = np.isfinite(ru) & ~eq
    lo = np.isfinite(rl) & ~eq
    A_ub = sp.vstack([A[up], -A[lo]]) if (up.any() or lo.any()) else None
    b_ub = np.concatenate([ru[up], -rl[lo]]) if A_ub is not None else None
    A_eq = A[eq] if eq.any() else None
    b_eq = rl[eq] if eq.any() else None
    bounds = [(None if not np.isfinite(l) else l, None if not np.isfinite(u) else u) for l, u in zip(xl, xu)]
    if np.any(xl > xu):
        return False
    res = linprog(
        np.zeros(n), A_ub=A_ub, b_ub=b_ub, A_eq=A_eq, b_eq=b_eq, bounds=bounds, method="highs",
        options={"primal_feasibility_tolerance": tol, "presolve": True},
    )
    return res.status != 2


def kkt_residuals(P, q, A, rl, ru, xl, xu, x, lam, rc):
    """Unscaled KKT residuals of a candidate (x, lambda, r) in MOI sign convention.

    Returns dict(stationarity, primal, complementarity, dual_sign) -- all inf-norms.
    """
    Px = P @ x if P is not None else np.zeros_like(x)
    Ax = A @ x
    stat = np.max(np.abs(Px + q - A.T @ lam - rc), initial=0.0)
    prim = max(
        np.max(np.maximum(rl - Ax, 0.0), initial=0.0),
        np.max(np.maximum(Ax - ru, 0.0), initial=0.0),
        np.max(np.maximum(xl - x, 0.0), initial=0.0),
        np.max(np.maximum(x - xu, 0.0), initial=0.0),
    )
    lp, ln = np.maximum(lam, 0.0), np.minimum(lam, 0.0)
    rp, rn = np.maximum(rc, 0.0), np.minimum(rc, 0.0)

    def _c(mult, gap):
        g = np.where(np.isfinite(gap), gap, 0.0)
        bad = np.where(np.isfinite(gap), 0.0, np.abs(mult))  # multiplier on an infinite bound
        return max(np.max(np.abs(mult * g), initial=0.0), np.max(bad, initial=0.0))

    comp = max(_c(lp, Ax - rl), _c(ln, ru - Ax), _c(rp, x - xl), _c(rn, xu - x))
    return {"stationarity": stat, "primal": prim, "complementarity": comp}


def _inertia_dense(K, nv):
    """Number of negative eigenvalues of symmetric K via dense LDL^T (Bunch-Kaufman)."""
    _, D, _ = sla.ldl(K, lower=True, hermitian=True)
    n = D.shape[0]
    neg = zero = 0
    i = 0
    while i < n:
        if i + 1 < n and D[i + 1, i] != 0.0:
            w = np.linalg.eigvalsh(D[i : i + 2, i : i + 2])
            neg += int(np.sum(w < 0))
            zero += int(np.sum(w == 0))
            i += 2
        else:
            neg += int(D[i, i] < 0)
            zero += int(D[i, i] == 0)
            i += 1
    return neg, zero


def solve_qp(P, q, A, rl, ru, xl, xu, x0=None, tol=1e-10, max_iter=400, check_feasibility=True,
             dense_inertia_limit=700, verbose=False):
    """Primal-dual interior-point QP/LP solve (see module docstring)."""
    q = np.asarray(q, float)
    n = q.shape[0]
    A = sp.csr_matrix(A) if A is not None else sp.csr_matrix((0, n))
    m = A.shape[0]
    rl, ru = np.asarray(rl, float).copy(), np.asarray(ru, float).copy()
    xl, xu = np.asarray(xl, float).copy(), np.asarray(xu, float).copy()
    P = sp.csr_matrix(P) if P is not None else sp.csr_matrix((n, n))
    zeros = QpResult(INFEASIBLE, np.zeros(n), np.zeros(m), np.zeros(n))
    if np.any(xl > xu) or np.any(rl > ru):
        return zeros
    if check_feasibility and not is_feasible(A, rl, ru, xl, xu):
        return zeros
    # gradient-based objective scaling (Ipopt's nlp_scaling_method default: max |grad| <= 100)
    gmax = float(np.max(np.abs(q), initial=0.0))
    fscale = 100.0 / gmax if gmax > 100.0 else 1.0
    P_orig, q_orig = P, q
    P = P * fscale
    q = q * fscale

    # ---- eliminate fixed columns ------------------------------------------------
    fixed = xl == xu
    free = ~fixed
    xfix = np.where(fixed, xl, 0.0)
    nf = int(free.sum())
    Pf = P[free][:, free].tocsr()
    qf = q[free] + (P[free] @ xfix)
    Af = A[:, free].tocsr()
    shift = A @ xfix
    rlf, ruf = rl - shift, ru - shift
    lx, ux = xl[free], xu[free]

    # presolve: rows with no non-zero coefficient on a free column carry no
    # information (feasibility was certified above); drop them, dual = 0
    row_absmax = np.zeros(m)
    if Af.nnz:
        np.maximum.at(row_absmax, np.repeat(np.arange(m), np.diff(Af.indptr)), np.abs(Af.data))
    live = row_absmax > 0.0
    eq = live & np.isfinite(rlf) & (rlf == ruf)
    iq = live & ~eq & (np.isfinite(rlf) | np.isfinite(ruf))
    nE, nI = int(eq.sum()), int(iq.sum())
    AE, AI = Af[eq], Af[iq]
    nv = nf + nI
    l = np.concatenate([lx, rlf[iq]])
    u = np.concatenate([ux, ruf[iq]])
    hasl, hasu = np.isfinite(l), np.isfinite(u)
    # keep a (sub-tolerance) strict interior when inequalities pin a value, in the
    # spirit of Ipopt's bound_relax_factor but two orders below ``tol``
    relax = 0.01 * tol
    l = np.where(hasl, l - relax * np.maximum(1.0, np.abs(np.where(hasl, l, 0.0))), l)
    u = np.where(hasu, u + relax * np.maximum(1.0, np.abs(np.where(hasu, u, 0.0))), u)
    C = sp.bmat([[AE, None if nI == 0 else sp.csr_matrix((nE, nI))],
                 [AI, -sp.identity(nI)]], format="csr") if nI > 0 else sp.csr_matrix(AE)
    if C.shape[1] != nv:  # nI == 0 and nE == 0
        C = sp.csr_matrix((0, nv))
    d = np.concatenate([rlf[eq], np.zeros(nI)])
    mC = nE + nI
    W = sp.block_diag([Pf, sp.csr_matrix((nI, nI))], format="csr") if nI > 0 else Pf
    qv = np.concatenate([qf, np.zeros(nI)])
    CT = C.T.tocsr()

    # ---- starting point ------------------------------------------------------------
    v = np.zeros(nv)
    if x0 is not None:
        v[:nf] = np.asarray(x0, float)[free]
    v[nf:] = AI @ v[:nf]
    k1 = k2 = 1e-2
    both = hasl & hasu
    pl = np.where(both, np.minimum(k1 * np.maximum(1.0, np.abs(np.where(hasl, l, 0.0))), k2 * np.where(both, u - l, 0.0)),
                  k1 * np.maximum(1.0, np.abs(np.where(hasl, l, 0.0))))
    pu = np.where(both, np.minimum(k1 * np.maximum(1.0, np.abs(np.where(hasu, u, 0.0))), k2 * np.where(both, u - l, 0.0)),
                  k1 * np.maximum(1.0, np.abs(np.where(hasu, u, 0.0))))
    v = np.where(hasl, np.maximum(v, np.where(hasl, l, 0.0) + pl), v)
    v = np.where(hasu, np.minimum(v, np.where(hasu, u, 0.0) - pu), v)
    lam = np.zeros(mC)
    zl = np.where(hasl, 1.0, 0.0)
    zu = np.where(hasu, 1.0, 0.0)
    mu = 0.1
    nu = 1.0  # merit penalty
    dw_last = 0.0
    kap_eps, kap_mu, th_mu, tau_min = 10.0, 0.2, 1.5, 0.99
    lf = np.where(hasl, l, 0.0)
    uf = np.where(hasu, u, 0.0)
    use_dense = (nv + mC) <= dense_inertia_limit

    def errors(mu_):
        sl = np.where(hasl, v - lf, 1.0)
        su = np.where(hasu, uf - v, 1.0)
        grad = W @ v + qv
        rd = grad + CT @ lam - zl + zu
        rp = C @ v - d
        cl = np.where(hasl, sl * zl - mu_, 0.0)
        cu = np.where(hasu, su * zu - mu_, 0.0)
        smax = 100.0
        zsum = np.abs(zl).sum() + np.abs(zu).sum()
        sd = max(smax, (np.abs(lam).sum() + zsum) / max(1, mC + nv)) / smax
        sc = max(smax, zsum / max(1, nv)) / smax
        e = max(np.max(np.abs(rd), initial=0.0) / sd, np.max(np.abs(rp), initial=0.0),
                max(np.max(np.abs(cl), initial=0.0), np.max(np.abs(cu), initial=0.0)) / sc)
        return e, grad, rd, rp, sl, su

    status = ITERATION_LIMIT
    it = 0
    e0 = np.inf
    # Ipopt's termination levels (options tol = 1e-8, acceptable_tol = 1e-6; Waechter & Biegler section 2.1 and the
    # "acceptable point" heuristic of the implementation): the loop aims at the tighter ``tol`` of this oracle, but an
    # iterate that met Ipopt's own default tolerance is a solution Ipopt would have returned, so the best iterate is
    # remembered and returned when the tail of the iteration stalls (tiny steps of a marginally feasible subproblem at
    # a small trust region, where the merit line search cannot make progress at the 1e-10 level).
    best = None
    tiny_steps = 0
    for it in range(max_iter):
        e0, grad, rd, rp, sl, su = errors(0.0)
        if best is None or e0 < best[0]:
            best = (e0, v.copy(), lam.copy(), zl.copy(), zu.copy())
        if e0 <= tol:
            status = LOCALLY_SOLVED
            break
        if tiny_steps >= 8 and best[0] <= 1e-6:
            break
        emu = errors(mu)[0]
        while emu <= kap_eps * mu and mu > tol / 10.0:
            mu = max(tol / 10.0, min(kap_mu * mu, mu**th_mu))
            emu = errors(mu)[0]
        tau = max(tau_min, 1.0 - mu)
        Sig = np.where(hasl, zl / sl, 0.0) + np.where(hasu, zu / su, 0.0)
        gphi = grad - np.where(hasl, mu / sl, 0.0) + np.where(hasu, mu / su, 0.0)
        rhs = np.concatenate([-(gphi + CT @ lam), -rp])
        # ---- factorise with inertia / curvature correction --------------------
        dw = 0.0
        dc = 0.0
        tries = 0
        while True:
            H = (W + sp.diags(Sig + dw)).tocsr()
            K = sp.bmat([[H, CT], [C, -dc * sp.identity(mC)]], format="csc") if mC > 0 else H.tocsc()
            ok = True
            try:
                if use_dense:
                    Kd = K.toarray()
                    neg, zero = _inertia_dense(Kd, nv)
                    if zero > 0 and dc == 0.0:
                        dc = 1e-8 * mu**0.25
                        continue
                    ok = neg == mC and zero == 0
                    if ok:
                        sol = np.linalg.solve(Kd, rhs)
                else:
                    lu = spla.splu(K)
                    sol = lu.solve(rhs)
                    if not np.all(np.isfinite(sol)):
                        raise RuntimeError("singular")
                    dv_ = sol[:nv]
                    curv = dv_ @ (H @ dv_)
                    ok = curv >= 1e-11 * (dv_ @ dv_)
                    # one step of iterative refinement
                    if ok:
                        sol = sol + lu.solve(rhs - K @ sol)
            except (RuntimeError, np.linalg.LinAlgError):
                ok = False
                if dc == 0.0:
                    dc = 1e-8 * mu**0.25
            if ok:
                break
            tries += 1
            if tries > 40:
                return QpResult(NUMERICAL_ERROR, np.zeros(n), np.zeros(m), np.zeros(n), iters=it)
            if dw == 0.0:
                dw = 1e-4 if dw_last == 0.0 else max(1e-20, dw_last / 3.0)
            else:
                dw *= 100.0 if dw_last == 0.0 else 8.0
        if dw > 0:
            dw_last = dw
        dv, dlam = sol[:nv], sol[nv:]
        dzl = np.where(hasl, mu / sl - zl - (zl / sl) * dv, 0.0)
        dzu = np.where(hasu, mu / su - zu + (zu / su) * dv, 0.0)

        def ftb(val, dval, mask):
            neg = mask & (dval < 0)
            if not neg.any():
                return 1.0
            return min(1.0, float(np.min(-tau * val[neg] / dval[neg])))

        a_p = min(ftb(sl, dv, hasl), ftb(su, -dv, hasu))
        a_d = min(ftb(zl, dzl, hasl), ftb(zu, dzu, hasu))
        # ---- l1-merit backtracking -------------------------------------------------
        c1n = np.abs(rp).sum()
        dphi = gphi @ dv
        curv = dv @ ((W @ dv) + Sig * dv)
        if c1n > 0:
            nu_trial = (dphi + 0.5 * max(curv, 0.0)) / (0.7 * c1n)
            if nu < nu_trial:
                nu = nu_trial + 1.0

        def merit(vv):
            s1 = np.where(hasl, vv - lf, 1.0)
            s2 = np.where(hasu, uf - vv, 1.0)
            if np.any(s1 <= 0) or np.any(s2 <= 0):
                return np.inf
            return (0.5 * vv @ (W @ vv) + qv @ vv - mu * np.log(s1[hasl]).sum() - mu * np.log(s2[hasu]).sum()
                    + nu * np.abs(C @ vv - d).sum())

        m0 = merit(v)
        Dm = dphi - nu * c1n
        a = a_p
        accepted = False
        for _ in range(40):
            if merit(v + a * dv) <= m0 + 1e-8 * a * Dm + 10.0 * np.finfo(float).eps * abs(m0):
                accepted = True
                break
            a *= 0.5
        if not accepted:
            a = a_p  # tiny-step regime near machine precision: take the Newton step
        tiny_steps = tiny_steps + 1 if a < 1e-5 else 0
        v = v + a * dv
        lam = lam + a * dlam
        zl = zl + a_d * dzl
        zu = zu + a_d * dzu
        # Ipopt's multiplier safeguard (eq. 16)
        ks = 1e10
        sl = np.where(hasl, v - lf, 1.0)
        su = np.where(hasu, uf - v, 1.0)
        zl = np.where(hasl, np.clip(zl, mu / (ks * sl), ks * mu / sl), 0.0)
        zu = np.where(hasu, np.clip(zu, mu / (ks * su), ks * mu / su), 0.0)
        if verbose:
            print(f"  ipm {it:3d} mu={mu:.1e} e0={e0:.2e} a={a:.2e} ad={a_d:.2e} dw={dw:.1e} nu={nu:.1e}")
    else:
        e0 = errors(0.0)[0]
        if e0 <= tol:
            status = LOCALLY_SOLVED
    if status != LOCALLY_SOLVED and best is not None and best[0] <= 1e-6:
        e0, v, lam, zl, zu = best
        status = LOCALLY_SOLVED if e0 <= 1e-8 else "ALMOST_LOCALLY_SOLVED"

    # ---- map back to MOI conventions -------------------------------------------------
    x = xfix.copy()
    x[free] = v[:nf]
    # L = f + lam'(Cv-d) - zl'(v-l) - zu'(u-v)  =>  grad f = -C' lam + zl - zu
    row_dual = np.zeros(m)
    lamE, lamI = lam[:nE], lam[nE:]
    row_dual[eq] = -lamE
    row_dual[iq] = -lamI  # (slack stationarity: -(-lamI) ... = zl_s - zu_s) -> same value
    col_dual = np.zeros(n)
    col_dual[free] = zl[:nf] - zu[:nf]
    # reduced cost of eliminated (fixed) columns from stationarity
    row_dual /= fscale
    col_dual /= fscale
    P, q = P_orig, q_orig
    if fixed.any():
        g = P @ x + q - A.T @ row_dual
        col_dual[fixed] = g[fixed]
    obj = 0.5 * x @ (P @ x) + q @ x
    return QpResult(status, x, row_dual, col_dual, obj=obj, iters=it, kkt_error=e0)
