"""DEVELOPMENT TOOL ONLY -- second numpy prototype of the device QP algorithm.

Neither product nor oracle (nothing imports it).  Same role as proto_admm.py,
but with the structure the CUDA engine implements:

  outer proximal-point loop (only when P is indefinite: nu = 1.2*|lambda_min|)
    inner OSQP-style ADMM on the convexified QP  (P + nu I,  q - nu x_c)
      PCG (Jacobi) on K = P + (nu+sigma) I + diag(rho_b) + J' diag(rho_c) J
      every `check_every` iterations: residuals, infeasibility certificate,
      rho adaptation, and -- once the predicted active set is stable --
    polish on the TRUE (P, q): fixed columns eliminated, active rows by the
      method of multipliers; accepted only if it is a verified KKT point.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from proto_admm import ruiz, lambda_min, pcg, _ninf, _support


class Opts:
    rho0 = 0.1
    sigma = 1e-6
    alpha = 1.6
    eps_abs = 1e-7
    eps_rel = 1e-7
    eps_inf = 1e-8
    max_iter = 4000
    check_every = 25
    rho_eq_mult = 1e3
    rho_min = 1e-6
    rho_max = 1e6
    adapt_tol = 3.0
    ruiz_iters = 15
    cg_max = 300
    eig_iters = 60
    nu_mult = 0.0
    rb_full_mult = 1.5
    prox_inner_rel = 1e-3     # inner convergence (relative residual) before the prox centre moves
    polish_trigger = 5e-2     # relative residual below which polish may be attempted
    polish_rho = 1e4
    polish_outer = 20
    polish_tol = 1e-11
    feas_tol = 1e-9
    dual_tol = 1e-9
    verbose = False
    polish_direct = False
    cg_rel0 = 0.1
    exact_reduced = False
    rb_mult = 2.0
    nu_red_mult = 0.0


def admm_solve(P, q, J, rl, ru, xl, xu, warm=None, o: Opts = Opts()):
    n, m = q.shape[0], J.shape[0]
    P = sp.csr_matrix(P) if P is not None else sp.csr_matrix((n, n))
    J = sp.csr_matrix(J)
    D, E, c = ruiz(P, J, q, o.ruiz_iters)
    Ps = (sp.diags(D) @ P @ sp.diags(D) * c).tocsr()
    Js = (sp.diags(E) @ J @ sp.diags(D)).tocsr()
    JsT = Js.T.tocsr()
    qs = c * D * q
    rls, rus = E * rl, E * ru
    xls, xus = xl / D, xu / D
    dP = Ps.diagonal()
    J2T = JsT.multiply(JsT).tocsr()
    nu = max(0.0, -lambda_min(Ps, o.eig_iters))
    rb_floor = o.rb_full_mult * nu
    nu = o.nu_mult * nu if nu > 1e-9 else 0.0
    if o.exact_reduced:  # experiment: exact reduced-Hessian curvature (dense; prototype only)
        import scipy.linalg as sla
        eq_ = (E * rl) == (E * ru)
        Z = sla.null_space(Js[eq_].toarray())
        wmin = np.linalg.eigvalsh(Z.T @ Ps.toarray() @ Z).min() if Z.shape[1] else 0.0
        nu_red = max(0.0, -wmin)
        rb_floor = o.rb_mult * nu_red
        nu = o.nu_red_mult * nu_red

    eqc = rls == rus
    freec = ~np.isfinite(rls) & ~np.isfinite(rus)
    eqb = xls == xus
    freeb = ~np.isfinite(xls) & ~np.isfinite(xus)

    def rho_vec(rho):
        rc = np.where(eqc, rho * o.rho_eq_mult, np.where(freec, o.rho_min, rho))
        rb = np.where(eqb, rho * o.rho_eq_mult, np.where(freeb, o.rho_min, rho))
        return rc, np.maximum(rb, rb_floor)

    rho = o.rho0
    if warm is not None:
        x = warm["x"] / D
        yc = c * warm["yc"] / E
        yb = c * warm["yb"] * D
        rho = warm.get("rho", rho)
    else:
        x, yc, yb = np.zeros(n), np.zeros(m), np.zeros(n)
    rc, rb = rho_vec(rho)
    zc = np.clip(Js @ x, rls, rus)
    zb = np.clip(x, xls, xus)
    xc = x.copy()
    xt_prev = x.copy()
    sigma = o.sigma
    info = dict(nu=nu, rb_floor=rb_floor, cg_iters=0, polish_tries=0, polish_cg=0, outers=1, rho_updates=0)
    status = "MAX_ITER"
    cg_rel = 1e-3
    last_act = None
    tried_act = None
    k = 0
    rp = rd = np.inf
    while k < o.max_iter:
        k += 1
        shift = nu + sigma
        Kmul = lambda v: Ps @ v + (shift + rb) * v + JsT @ (rc * (Js @ v))
        Minv = 1.0 / (dP + shift + rb + J2T @ rc)
        qc = qs - nu * xc
        rhs = sigma * x - qc + JsT @ (rc * zc - yc) + (rb * zb - yb)
        xt, it, rn, neg = pcg(Kmul, Minv, rhs, xt_prev, 1e-14 * np.sqrt(rhs @ rhs), o.cg_max, rel0=o.cg_rel0)
        xt_prev = xt
        info["cg_iters"] += it
        if neg:  # eigenvalue estimate too optimistic
            nu = max(2 * nu, 1e-3)
            info["nu"] = nu
            continue
        ztc = Js @ xt
        xn = o.alpha * xt + (1 - o.alpha) * x
        zc_r = o.alpha * ztc + (1 - o.alpha) * zc
        zb_r = o.alpha * xt + (1 - o.alpha) * zb
        zcn = np.clip(zc_r + yc / rc, rls, rus)
        zbn = np.clip(zb_r + yb / rb, xls, xus)
        ycn = yc + rc * (zc_r - zcn)
        ybn = yb + rb * (zb_r - zbn)
        dy_c, dy_b = ycn - yc, ybn - yb
        x, zc, zb, yc, yb = xn, zcn, zbn, ycn, ybn
        if k % o.check_every:
            continue
        Ax = Js @ x
        Px = Ps @ x + nu * x
        ATy = JsT @ yc + yb
        rp = max(_ninf((Ax - zc) / E), _ninf((x - zb) * D))
        rd = _ninf((Px + qc + ATy) / D) / c
        np_ = max(_ninf(Ax / E), _ninf(zc / E), _ninf(x * D), _ninf(zb * D))
        nd_ = max(_ninf(Px / D), _ninf(ATy / D), _ninf(qc / D)) / c
        relp, reld = rp / max(np_, 1e-30), rd / max(nd_, 1e-30)
        if o.verbose:
            print(f"  admm {k:5d} rp={rp:.2e}({relp:.1e}) rd={rd:.2e}({reld:.1e}) rho={rho:.2e} nu={nu:.2e} cg={info['cg_iters']}")
        # infeasibility certificate
        dyc_u, dyb_u = E * dy_c / c, dy_b / D / c
        ndy = max(_ninf(dyc_u), _ninf(dyb_u))
        if ndy > 1e-30:
            ATdy = _ninf((JsT @ dy_c + dy_b) / D) / c
            if ATdy <= o.eps_inf * ndy and (_support(dyc_u, rl, ru) + _support(dyb_u, xl, xu)) <= -o.eps_inf * ndy:
                status = "PRIMAL_INFEASIBLE"
                break
        # active-set prediction and polish
        act = _active(x, zc, yc, yb, rls, rus, xls, xus)
        key = tuple(a.tobytes() for a in act)
        if max(relp, reld) < o.polish_trigger and key == last_act and key != tried_act:
            tried_act = key
            info["polish_tries"] += 1
            res = polish(Ps, qs, Js, JsT, rls, rus, xls, xus, x, act, yc, dP, J2T, o)
            if res is not None:
                x2, yc2, yb2, pinfo = res
                info["polish_cg"] += pinfo["cg"]
                if pinfo["verified"]:
                    x, yc, yb = x2, yc2, yb2
                    status = "SOLVED"
                    info["polished"] = True
                    info["polish"] = pinfo
                    break
        last_act = key
        if nu == 0.0 and rp <= o.eps_abs + o.eps_rel * np_ and rd <= o.eps_abs + o.eps_rel * nd_:
            status = "SOLVED"
            info["polished"] = False
            break
        if nu > 0.0 and max(relp, reld) < o.prox_inner_rel:
            move = _ninf((x - xc) * D)
            xc = x.copy()
            info["outers"] += 1
            if move <= o.eps_abs and rp <= o.eps_abs + o.eps_rel * np_:
                status = "SOLVED"
                info["polished"] = False
                break
            continue
        # rho adaptation
        new = rho * np.sqrt(relp / max(reld, 1e-30))
        new = min(max(new, o.rho_min), o.rho_max)
        if new > rho * o.adapt_tol or new < rho / o.adapt_tol:
            rho = new
            rc, rb = rho_vec(rho)
            info["rho_updates"] += 1
        cg_rel = min(cg_rel, max(0.1 * min(relp, reld), 1e-9))
    info.update(admm_iters=k, rho=rho, rp=rp, rd=rd)
    return {"status": status, "x": D * x, "yc": E * yc / c, "yb": yb / D / c, "rho": rho, "info": info}


def _active(x, zc, yc, yb, rls, rus, xls, xus):
    lowc = ((zc - rls < -yc) | (rls == rus))
    upc = (rus - zc < yc) & ~lowc
    lowb = ((x - xls < -yb) | (xls == xus))
    upb = (xus - x < yb) & ~lowb
    return lowc, upc, lowb, upb


def polish(Ps, qs, Js, JsT, rls, rus, xls, xus, x, act, yc, dP, J2T, o):
    lowc, upc, lowb, upb = act
    actc = lowc | upc
    bc = np.where(lowc, rls, np.where(upc, rus, 0.0))
    free = ~(lowb | upb)
    xfix = np.where(lowb, xls, np.where(upb, xus, 0.0))
    mask = free.astype(float)
    w = actc.astype(float)
    rhoP, sig = o.polish_rho, 1e-9
    y = np.where(actc, yc, 0.0)
    xp = np.where(free, x, xfix)
    Kmul = lambda v: Ps @ v + sig * v + rhoP * (JsT @ (w * (Js @ v)))
    Minv = 1.0 / np.maximum(dP + sig + rhoP * (J2T @ w), 1e-12)
    tot = 0
    r = np.zeros_like(w)
    if o.polish_direct:
        import scipy.sparse.linalg as spla
        fi = np.nonzero(free)[0]
        Kf = (Ps + sig * sp.identity(Ps.shape[0]) + rhoP * (JsT @ sp.diags(w) @ Js)).tocsc()
        Kff = Kf[fi][:, fi].tocsc()
        lu = spla.splu(Kff)
    for it in range(o.polish_outer):
        rhs = sig * xp - qs + JsT @ (w * (rhoP * bc - y))
        if o.polish_direct:
            xc = np.where(free, 0.0, xfix)
            rf = (rhs - Kf @ xc)[fi]
            xf = lu.solve(rf)
            xf += lu.solve(rf - Kff @ xf)
            xt = xc.copy(); xt[fi] = xf; cg = 1; neg = False
        else:
            xt, cg, rn, neg = pcg(Kmul, Minv, rhs, xp, max(1e-13 * np.sqrt(rhs @ rhs), 1e-300), 3000, mask=mask)
        tot += cg
        if neg:
            return None
        xp = np.where(free, xt, xfix)
        r = w * (Js @ xp - bc)
        y = y + rhoP * r
        if _ninf(r) < o.polish_tol:
            break
    g = Ps @ xp + qs + JsT @ y
    ybn = np.where(free, 0.0, -g)
    stat = _ninf(g * mask)
    Ax = Js @ xp
    pf = max(_ninf(np.maximum(rls - Ax, 0)), _ninf(np.maximum(Ax - rus, 0)), _ninf(np.maximum(xls - xp, 0)), _ninf(np.maximum(xp - xus, 0)))
    # dual signs (OSQP convention: y>0 at upper, y<0 at lower); equalities are free
    eqc = rls == rus
    eqb = xls == xus
    ds = max(_ninf(np.maximum(y, 0) * (lowc & ~eqc)), _ninf(np.minimum(y, 0) * upc),
             _ninf(np.maximum(ybn, 0) * (lowb & ~eqb)), _ninf(np.minimum(ybn, 0) * upb))
    ymag = max(1.0, _ninf(y), _ninf(ybn))
    ok = pf <= o.feas_tol and ds <= o.dual_tol * ymag and _ninf(r) <= 100 * o.polish_tol and stat <= 1e-9 * ymag
    return xp, y, ybn, {"outer": it + 1, "cg": tot, "res": _ninf(r), "pf": pf, "ds": ds, "stat": stat, "verified": bool(ok)}
