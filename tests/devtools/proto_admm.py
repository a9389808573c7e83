"""DEVELOPMENT TOOL ONLY -- numpy prototype of the device QP algorithm.

Neither product nor oracle: nothing under sqpsolver.jl_b200/, tests/ or bench.py
imports it.  It exists because the build container has no GPU: the ADMM + PCG +
polish algorithm that csrc/ implements is tuned here on recorded QP sequences
(iteration counts, rho/sigma rules, polish strategy) before GPU minutes are spent.
The arithmetic mirrors the kernels one-to-one (same scaling, same update order),
but uses scipy CSR products instead of the hand-written SpMV.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

INF = np.inf


class Opts:
    rho0 = 0.1
    sigma = 1e-6
    alpha = 1.6
    eps_abs = 1e-6
    eps_rel = 1e-6
    eps_inf = 1e-7
    max_iter = 20000
    check_every = 10
    rho_eq_mult = 1e3
    rho_min = 1e-6
    rho_max = 1e6
    adapt_every = 50
    adapt_tol = 2.0
    ruiz_iters = 15
    cg_max = 500
    cg_tol_min = 1e-10
    polish = True
    eig_iters = 60
    nc_rho_mult = 0.0
    nc_sigma_mult = 0.0
    polish_rho = 1e4
    polish_outer = 30
    polish_tol = 1e-11
    verbose = False


def ruiz(P, J, q, iters):
    n = P.shape[0]
    m = J.shape[0]
    D = np.ones(n)
    E = np.ones(m)
    c = 1.0
    Pa = abs(P).tocsr()
    Ja = abs(J).tocsr()
    JTa = Ja.T.tocsr()
    for _ in range(iters):
        # column norms of the scaled KKT matrix [cDPD  DJ'E; EJD 0]
        pn = c * D * _rowmax(Pa, D)
        jn = D * _rowmax(JTa, E)
        cn = np.maximum(pn, jn)
        rn = E * _rowmax(Ja, D)
        dD = 1.0 / np.sqrt(np.where(cn > 1e-4, np.minimum(cn, 1e4), 1.0))
        dE = 1.0 / np.sqrt(np.where(rn > 1e-4, np.minimum(rn, 1e4), 1.0))
        D *= dD
        E *= dE
        # cost scaling
        pn = c * D * _rowmax(Pa, D)
        qn = np.max(np.abs(c * D * q), initial=0.0)
        avg = pn.mean() if n else 1.0
        g = max(avg, qn)
        g = 1.0 / (g if 1e-4 < g else 1.0)
        g = min(max(g, 1e-4), 1e4)
        c *= g
    return D, E, c


def _rowmax(Mabs, colscale):
    """max_j |M_ij| * colscale_j for each row i."""
    out = np.zeros(Mabs.shape[0])
    if Mabs.nnz:
        vals = Mabs.data * colscale[Mabs.indices]
        rows = np.repeat(np.arange(Mabs.shape[0]), np.diff(Mabs.indptr))
        np.maximum.at(out, rows, vals)
    return out


def lambda_min(Ps, iters):
    """Smallest eigenvalue of symmetric Ps: power iteration on (bound*I - Ps)."""
    n = Ps.shape[0]
    if Ps.nnz == 0:
        return 0.0
    bound = float(abs(Ps).sum(axis=1).max())  # Gershgorin
    v = np.cos(np.arange(n) * 0.7 + 0.3)  # deterministic start
    v /= np.linalg.norm(v)
    lam = 0.0
    for _ in range(iters):
        w = bound * v - Ps @ v
        lam = v @ w
        nw = np.linalg.norm(w)
        if nw == 0:
            break
        v = w / nw
    return bound - lam


def pcg(Kmul, Minv, b, x0, tol_abs, maxit, mask=None, rel0=0.0):
    x = x0.copy()
    r = b - Kmul(x)
    if mask is not None:
        r = r * mask
    tol_abs = max(tol_abs, rel0 * np.sqrt(r @ r))
    z = Minv * r
    p = z.copy()
    rz = r @ z
    it = 0
    negcurv = False
    rn = np.sqrt(r @ r)
    while rn > tol_abs and it < maxit:
        Kp = Kmul(p)
        if mask is not None:
            Kp = Kp * mask
        pKp = p @ Kp
        if pKp <= 0:
            negcurv = True
            break
        a = rz / pKp
        x += a * p
        r -= a * Kp
        z = Minv * r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
        rn = np.sqrt(r @ r)
        it += 1
    return x, it, rn, negcurv


def admm_solve(P, q, J, rl, ru, xl, xu, warm=None, o: Opts = Opts()):
    """OSQP-form ADMM on  min 1/2x'Px+q'x, rl<=Jx<=ru, xl<=x<=xu.  Returns dict."""
    n = q.shape[0]
    m = J.shape[0]
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

    def rho_vec(rho):
        rc = np.full(m, rho)
        rc[rls == rus] = rho * o.rho_eq_mult
        rc[~np.isfinite(rls) & ~np.isfinite(rus)] = o.rho_min
        rb = np.full(n, rho)
        rb[xls == xus] = rho * o.rho_eq_mult
        rb[~np.isfinite(xls) & ~np.isfinite(xus)] = o.rho_min
        return np.clip(rc, o.rho_min, o.rho_max * o.rho_eq_mult), np.clip(rb, o.rho_min, o.rho_max * o.rho_eq_mult)

    rho = o.rho0
    sigma = o.sigma
    # nonconvexity: most negative eigenvalue of the scaled P by shifted power iteration
    nu = max(0.0, -lambda_min(Ps, o.eig_iters))
    info_nu = nu
    rho_b_floor = o.nc_rho_mult * nu
    sigma = max(sigma, o.nc_sigma_mult * nu)
    _rho_vec0 = rho_vec

    def rho_vec(rho):
        rc_, rb_ = _rho_vec0(rho)
        return rc_, np.maximum(rb_, rho_b_floor)

    rc, rb = rho_vec(rho)
    if warm is not None:
        x = warm["x"] / D
        yc = c * warm["yc"] / E
        yb = c * warm["yb"] * D
        rho = warm.get("rho", rho)
        rc, rb = rho_vec(rho)
    else:
        x = np.zeros(n)
        yc = np.zeros(m)
        yb = np.zeros(n)
    zc = np.clip(Js @ x, rls, rus)
    zb = np.clip(x, xls, xus)

    total_cg = 0
    status = "MAX_ITER"
    info = {}
    cg_tol_rel = 1e-2
    n_rho = 0
    dy_c = np.zeros(m)
    dy_b = np.zeros(n)
    sig_bumps = 0
    k = 0
    rp = rd = np.inf
    while k < o.max_iter:
        k += 1
        Kmul = lambda v: Ps @ v + (sigma + rb) * v + JsT @ (rc * (Js @ v))
        Minv = 1.0 / (dP + sigma + rb + J2T @ rc)
        rhs = sigma * x - qs + JsT @ (rc * zc - yc) + (rb * zb - yb)
        tol_abs = max(cg_tol_rel * np.sqrt(rhs @ rhs), 1e-15)
        xt, it, rn, neg = pcg(Kmul, Minv, rhs, x, tol_abs, o.cg_max)
        total_cg += it
        if neg:
            sigma = max(sigma * 10.0, 1e-4)
            sig_bumps += 1
            if sig_bumps > 30:
                status = "NONCONVEX"
                break
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
        if k % o.check_every == 0 or k == 1:
            Ax = Js @ x
            Px = Ps @ x
            ATy = JsT @ yc + yb
            rp = max(_ninf((Ax - zc) / E), _ninf((x - zb) * D))
            rd = _ninf((Px + qs + ATy) / D) / c
            np_ = max(_ninf(Ax / E), _ninf(zc / E), _ninf(x * D), _ninf(zb * D))
            nd_ = max(_ninf(Px / D), _ninf(ATy / D), _ninf(qs / D)) / c
            eps_p = o.eps_abs + o.eps_rel * np_
            eps_d = o.eps_abs + o.eps_rel * nd_
            if o.verbose:
                print(f"  admm {k:5d} rp={rp:.2e} rd={rd:.2e} rho={rho:.2e} sig={sigma:.1e} cg={total_cg}")
            if rp <= eps_p and rd <= eps_d:
                status = "SOLVED"
                break
            # primal infeasibility certificate (OSQP sec. 3.4) on unscaled dy
            dyc_u, dyb_u = E * dy_c / c, dy_b / D / c
            ndy = max(_ninf(dyc_u), _ninf(dyb_u))
            if ndy > 1e-30:
                ATdy = _ninf((JsT @ dy_c + dy_b) / D) / c
                sup = (_support(dyc_u, rl, ru) + _support(dyb_u, xl, xu))
                if ATdy <= o.eps_inf * ndy and sup <= -o.eps_inf * ndy:
                    status = "PRIMAL_INFEASIBLE"
                    break
            # adaptive rho
            if k % o.adapt_every == 0:
                sp_ = rp / max(np_, 1e-30)
                sd_ = rd / max(nd_, 1e-30)
                new = rho * np.sqrt(sp_ / max(sd_, 1e-30))
                new = min(max(new, o.rho_min), o.rho_max)
                if new > rho * o.adapt_tol or new < rho / o.adapt_tol:
                    rho = new
                    rc, rb = rho_vec(rho)
                    n_rho += 1
            # CG tolerance follows the residuals
            cg_tol_rel = min(cg_tol_rel, max(0.15 * np.sqrt(max(rp, 1e-30) * max(rd, 1e-30)) / max(np.sqrt(rhs @ rhs), 1e-30), o.cg_tol_min))
            cg_tol_rel = max(cg_tol_rel, o.cg_tol_min)
    info.update(nu=info_nu, admm_iters=k, cg_iters=total_cg, rho=rho, sigma=sigma, rho_updates=n_rho, rp=rp, rd=rd)
    polished = False
    if status == "SOLVED" and o.polish:
        res = polish(Ps, qs, Js, JsT, rls, rus, xls, xus, x, zc, yc, yb, dP, J2T, o)
        if res is not None:
            x2, yc2, yb2, pinfo = res
            # accept if the unscaled residuals improved
            Ax = Js @ x2
            rp2 = max(_ninf(np.maximum(rls - Ax, 0) / E), _ninf(np.maximum(Ax - rus, 0) / E),
                      _ninf(np.maximum(xls - x2, 0) * D), _ninf(np.maximum(x2 - xus, 0) * D))
            rd2 = _ninf((Ps @ x2 + qs + JsT @ yc2 + yb2) / D) / c
            info.update(polish=pinfo, rp_polish=rp2, rd_polish=rd2)
            if max(rp2, rd2) < max(rp, rd):
                x, yc, yb = x2, yc2, yb2
                polished = True
    info["polished"] = polished
    return {
        "status": status, "x": D * x, "yc": E * yc / c, "yb": yb / D / c, "rho": rho, "info": info,
    }


def polish(Ps, qs, Js, JsT, rls, rus, xls, xus, x, zc, yc, yb, dP, J2T, o):
    """Active-set refinement: fixed columns eliminated, active rows by method of multipliers."""
    n, m = x.shape[0], zc.shape[0]
    lowc = (zc - rls < -yc) | (rls == rus)
    upc = (rus - zc < yc) & ~lowc
    lowb = (x - xls < -yb) | (xls == xus)
    upb = (xus - x < yb) & ~lowb
    actc = lowc | upc
    bc = np.where(lowc, rls, np.where(upc, rus, 0.0))
    free = ~(lowb | upb)
    xfix = np.where(lowb, xls, np.where(upb, xus, 0.0))
    mask = free.astype(float)
    w = actc.astype(float)
    rhoP = o.polish_rho
    sig = 1e-9
    y = np.where(actc, yc, 0.0)
    xp = np.where(free, x, xfix)
    Kmul = lambda v: Ps @ v + sig * v + rhoP * (JsT @ (w * (Js @ v)))
    Minv = 1.0 / (dP + sig + rhoP * (J2T @ w))
    tot = 0
    res_hist = []
    for it in range(o.polish_outer):
        # solve for free part with fixed part held at bounds:  K_ff x_f = rhs_f - K_fc x_c
        rhs = sig * xp - qs + JsT @ (w * (rhoP * bc - y))
        xt, cg, rn, neg = pcg(Kmul, Minv, rhs, xp, max(1e-13 * np.sqrt(rhs @ rhs), 1e-300), 2000, mask=mask)
        tot += cg
        if neg:
            return None
        xp = np.where(free, xt, xfix)
        r = w * (Js @ xp - bc)
        y = y + rhoP * r
        res_hist.append(_ninf(r))
        if _ninf(r) < o.polish_tol:
            break
    ybn = -(Ps @ xp + qs + JsT @ y)
    ybn = np.where(free, 0.0, ybn)
    return xp, y, ybn, {"outer": it + 1, "cg": tot, "res": res_hist[-1] if res_hist else 0.0, "nact": int(actc.sum()), "nfix": int((~free).sum())}


def _ninf(v):
    return float(np.max(np.abs(v), initial=0.0))


def _support(dy, l, u):
    """u'(dy)+ + l'(dy)-  with infinite bounds contributing only if dy has the wrong sign."""
    pos = np.maximum(dy, 0.0)
    neg = np.minimum(dy, 0.0)
    with np.errstate(invalid="ignore"):
        a = np.where(pos > 0, u * pos, 0.0)
        b = np.where(neg < 0, l * neg, 0.0)
    return float(np.sum(a) + np.sum(b))
