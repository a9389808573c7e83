"""DEVELOPMENT TOOL ONLY -- numpy prototype of the device interior-point path (method = IPM).

Neither product nor oracle (nothing imports it).  Regularised primal-dual interior point
method with a CONDENSED, always-SPD Newton system that has exactly the structure of the ADMM
matrix the engine already applies:

    K = P + diag(rho_p + w_box) + J' diag(w_row) J,      w_row = 1/delta on equality rows,
                                                          z/(s + delta z) summed over the finite
                                                          sides on inequality rows
solved by sparse Cholesky (on device: batched, shared symbolic analysis; here: scipy splu).
Indefinite P -> factorisation breaks down -> rho_p is raised (Ipopt-style inertia correction).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from proto_admm import ruiz, _ninf


class Opts:
    eps = 1e-9
    max_iter = 80
    delta0 = 1e-6
    rho0 = 1e-8
    delta_min = 1e-11
    tau = 0.995
    ruiz_iters = 15
    verbose = False
    use_sym = False
    monotone = True
    mu0 = 1.0
    kappa_eps = 10.0
    rho_dec = 3.0      # decay of the inertia-correction shift per iteration
    rho_bump = 10.0    # growth on a failed factorisation
    rho_hold = 0       # iterations the shift is held after a failed factorisation
    rho_floor_frac = 0.0  # never decay below this fraction of the last shift that was NEEDED
    mu_force = 0       # force a barrier decrease after this many iterations without one (0 = never)
    loqo = False       # LOQO-style adaptive barrier parameter instead of the monotone rule
    split_step = False # separate primal and dual step lengths (Ipopt) instead of one common step
    sigma_rule = 0.0   # > 0: barrier target = sigma * current average complementarity, every iteration
    kappa_mu = 0.2     # linear / superlinear decrease of the barrier parameter (Ipopt: 0.2, 1.5)
    theta_mu = 1.5
    refine = True      # one step of iterative refinement of the Newton solve against the assembled K
    factor = None      # optional callable K -> (solver with .solve(rhs), ok): e.g. a modified factorisation


class _SymLU:
    """Linear solves through the host-executed device index programs (tests/support)."""
    def __init__(self, Ps, Js, dvec, wvec):
        import ctypes as C, sys
        sys.path.insert(0, '/root/repo/tests')
        self.lib = C.CDLL('/root/repo/tests/support/libsymcheck.so')
        self.args = (Ps, Js, dvec, wvec)
        r = self.solve(np.ones(Ps.shape[0]))
        self.ok = r is not None
    def solve(self, rhs):
        import ctypes as C
        Ps, Js, dvec, wvec = self.args
        ip=C.POINTER(C.c_int32); dp=C.POINTER(C.c_double)
        a=lambda v,t: np.ascontiguousarray(v,dtype=t)
        n=Ps.shape[0]; x=np.zeros(n)
        A=[a(Js.indptr,np.int32),a(Js.indices,np.int32),a(Js.data,np.float64),a(Ps.indptr,np.int32),a(Ps.indices,np.int32),a(Ps.data,np.float64),a(dvec,np.float64),a(wvec,np.float64),a(rhs,np.float64)]
        rc=self.lib.symcheck_solve(n,Js.shape[0],A[0].ctypes.data_as(ip),A[1].ctypes.data_as(ip),A[2].ctypes.data_as(dp),A[3].ctypes.data_as(ip),A[4].ctypes.data_as(ip),A[5].ctypes.data_as(dp),A[6].ctypes.data_as(dp),A[7].ctypes.data_as(dp),A[8].ctypes.data_as(dp),x.ctypes.data_as(dp),None)
        return x if rc==0 else None


def chol_solve(K, rhs):
    """Return (solve, ok).  ok=False if K is not positive definite (negative/zero pivot)."""
    try:
        lu = spla.splu(K.tocsc(), permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0, options=dict(SymmetricMode=True))
    except RuntimeError:
        return None, False
    d = lu.U.diagonal()
    if np.any(d <= 0) or not np.all(np.isfinite(d)):
        return None, False
    return lu, True


def ipm_solve(P, q, J, rl, ru, xl, xu, o: Opts = Opts()):
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
    eqr = rls == rus
    eqx = xls == xus
    # finite inequality sides
    ru_f = np.isfinite(rus) & ~eqr
    rl_f = np.isfinite(rls) & ~eqr
    xu_f = np.isfinite(xus) & ~eqx
    xl_f = np.isfinite(xls) & ~eqx
    nin = int(ru_f.sum() + rl_f.sum() + xu_f.sum() + xl_f.sum())
    x = np.clip(np.zeros(n), np.where(np.isfinite(xls), xls, -np.inf), np.where(np.isfinite(xus), xus, np.inf))
    x = np.where(eqx, xls, x)
    Ax = Js @ x
    # slacks/duals per side (arrays full length, masked)
    def init_s(gap):
        return np.maximum(gap, 1.0)
    s_ru = np.where(ru_f, init_s(rus - Ax), 1.0); z_ru = np.where(ru_f, 1.0, 0.0)
    s_rl = np.where(rl_f, init_s(Ax - rls), 1.0); z_rl = np.where(rl_f, 1.0, 0.0)
    s_xu = np.where(xu_f, init_s(xus - x), 1.0); z_xu = np.where(xu_f, 1.0, 0.0)
    s_xl = np.where(xl_f, init_s(x - xls), 1.0); z_xl = np.where(xl_f, 1.0, 0.0)
    y = np.zeros(m)      # equality-row duals
    yx = np.zeros(n)     # fixed-variable duals
    delta, rho_p = o.delta0, o.rho0
    rho_last = 0.0
    mu_t = o.mu0
    status = "MAX_ITER"
    nfact = 0
    hold = 0
    rho_need = 0.0
    for it in range(o.max_iter):
        Ax = Js @ x
        Px = Ps @ x
        lam_row = np.where(eqr, y, z_ru - z_rl)          # multiplier of each row in OSQP sign (>0 upper)
        lam_box = np.where(eqx, yx, z_xu - z_xl)
        r_x = Px + qs + JsT @ lam_row + lam_box
        r_eq = np.where(eqr, Ax - rls, 0.0)
        r_eqx = np.where(eqx, x - xls, 0.0)
        r_ru = np.where(ru_f, Ax + s_ru - rus, 0.0)
        r_rl = np.where(rl_f, -Ax + s_rl + rls, 0.0)
        r_xu = np.where(xu_f, x + s_xu - xus, 0.0)
        r_xl = np.where(xl_f, -x + s_xl + xls, 0.0)
        mu = (s_ru @ z_ru + s_rl @ z_rl + s_xu @ z_xu + s_xl @ z_xl) / max(nin, 1)
        rp = max(_ninf(r_eq / E), _ninf(r_ru / E), _ninf(r_rl / E), _ninf(r_eqx * D), _ninf(r_xu * D), _ninf(r_xl * D))
        rd = _ninf(r_x / D) / c
        if o.verbose:
            print(f"  ipm {it:3d} rp={rp:.2e} rd={rd:.2e} mu={mu:.2e} delta={delta:.1e} rho={rho_p:.1e} nfact={nfact}")
        scale_p = max(1.0, _ninf(Ax / E), _ninf(x * D))
        scale_d = max(1.0, _ninf(Px / D) / c, _ninf(qs / D) / c, _ninf((JsT @ lam_row) / D) / c)
        if rp <= o.eps * scale_p and rd <= o.eps * scale_d and mu <= o.eps * max(1.0, scale_d) * c:
            status = "SOLVED"
            break
        # weights
        d_ru, d_rl = s_ru + delta * z_ru, s_rl + delta * z_rl
        d_xu, d_xl = s_xu + delta * z_xu, s_xl + delta * z_xl
        w_row = np.where(eqr, 1.0 / delta, np.where(ru_f, z_ru / d_ru, 0.0) + np.where(rl_f, z_rl / d_rl, 0.0))
        w_box = np.where(eqx, 1.0 / delta, np.where(xu_f, z_xu / d_xu, 0.0) + np.where(xl_f, z_xl / d_xl, 0.0))
        while True:
            K = (Ps + sp.diags(rho_p + w_box) + JsT @ sp.diags(w_row) @ Js).tocsc()
            if o.factor is not None:
                lu, ok = o.factor(K)
            elif o.use_sym:
                lu = _SymLU(Ps, Js, rho_p + w_box, w_row); ok = lu.ok
            else:
                lu, ok = chol_solve(K, None)
            nfact += 1
            if ok:
                break
            hold = o.rho_hold
            rho_p = max(o.rho_bump * rho_p, 1e-4 if rho_last == 0.0 else rho_last / 3.0, 1e-6)
            rho_need = rho_p
            if rho_p > 1e8:
                return {"status": "NUMERICAL", "x": D * x, "info": dict(iters=it, nfact=nfact)}
        if rho_p > o.rho0 * 10:
            rho_last = rho_p

        def solve_dir(rc_ru, rc_rl, rc_xu, rc_xl):
            # rhs = -r_x - A_E' r_eq/delta - sum_k g_k (rc_k + z_k r_k)/d_k     (g = +J for upper, -J for lower)
            t_row = np.where(eqr, r_eq / delta, np.where(ru_f, (rc_ru + z_ru * r_ru) / d_ru, 0.0) - np.where(rl_f, (rc_rl + z_rl * r_rl) / d_rl, 0.0))
            t_box = np.where(eqx, r_eqx / delta, np.where(xu_f, (rc_xu + z_xu * r_xu) / d_xu, 0.0) - np.where(xl_f, (rc_xl + z_xl * r_xl) / d_xl, 0.0))
            rhs = -r_x - JsT @ t_row - t_box
            dx = lu.solve(rhs)
            if o.refine:
                dx += lu.solve(rhs - K @ dx)
            Jdx = Js @ dx
            dz_ru = np.where(ru_f, (rc_ru + z_ru * (r_ru + Jdx)) / d_ru, 0.0)
            dz_rl = np.where(rl_f, (rc_rl + z_rl * (r_rl - Jdx)) / d_rl, 0.0)
            dz_xu = np.where(xu_f, (rc_xu + z_xu * (r_xu + dx)) / d_xu, 0.0)
            dz_xl = np.where(xl_f, (rc_xl + z_xl * (r_xl - dx)) / d_xl, 0.0)
            ds_ru = np.where(ru_f, -r_ru - Jdx + delta * dz_ru, 0.0)
            ds_rl = np.where(rl_f, -r_rl + Jdx + delta * dz_rl, 0.0)
            ds_xu = np.where(xu_f, -r_xu - dx + delta * dz_xu, 0.0)
            ds_xl = np.where(xl_f, -r_xl + dx + delta * dz_xl, 0.0)
            dy = np.where(eqr, (Jdx + r_eq) / delta, 0.0)
            dyx = np.where(eqx, (dx + r_eqx) / delta, 0.0)
            return dx, (dz_ru, dz_rl, dz_xu, dz_xl), (ds_ru, ds_rl, ds_xu, ds_xl), dy, dyx

        def max_step(vs, dvs, masks):
            a = 1.0
            for v, dv, mk in zip(vs, dvs, masks):
                neg = mk & (dv < 0)
                if neg.any():
                    a = min(a, float(np.min(-v[neg] / dv[neg])))
            return a

        S = (s_ru, s_rl, s_xu, s_xl); Z = (z_ru, z_rl, z_xu, z_xl); MK = (ru_f, rl_f, xu_f, xl_f)
        if o.monotone:
            # barrier subproblem error; decrease mu_t when it is solved well enough
            mu_before = mu_t
            while True:
                comp = max(_ninf((s * z - mu_t) * mk) for s, z, mk in zip(S, Z, MK))
                e_mu = max(rd * c if False else _ninf(r_x), _ninf(r_eq), _ninf(r_ru), _ninf(r_rl), _ninf(r_eqx), _ninf(r_xu), _ninf(r_xl), comp)
                if e_mu <= o.kappa_eps * mu_t and mu_t > 1e-13:
                    mu_t = max(1e-13, min(o.kappa_mu * mu_t, mu_t ** o.theta_mu))
                else:
                    break
            if mu_t < mu_before:
                stall = 0
            else:
                stall = locals().get('stall', 0) + 1
                if o.mu_force and stall >= o.mu_force and mu_t > 1e-13:
                    mu_t = max(1e-13, 0.2 * mu_t)
                    stall = 0
            if o.loqo:
                prods = np.concatenate([(s * z)[mk] for s, z, mk in zip(S, Z, MK)])
                avg = prods.mean(); xi = prods.min() / avg
                sig = 0.1 * min(0.05 * (1 - xi) / max(xi, 1e-12), 2.0) ** 3
                mu_t = max(1e-13, sig * avg)
            if o.sigma_rule > 0.0:
                mu_t = max(1e-13, o.sigma_rule * mu)
            rc = [mu_t - s * z for s, z in zip(S, Z)]
            dx, dZ, dS, dy, dyx = solve_dir(*rc)
            tau = max(o.tau, 1.0 - mu_t)
            a = min(1.0, tau * min(max_step(S, dS, MK), max_step(Z, dZ, MK)))
            ad = a
            if o.split_step:
                a = min(1.0, tau * max_step(S, dS, MK)); ad = min(1.0, tau * max_step(Z, dZ, MK))
            x = x + a * dx; y = y + ad * dy; yx = yx + ad * dyx
            s_ru, s_rl, s_xu, s_xl = [s + a * ds for s, ds in zip(S, dS)]
            z_ru, z_rl, z_xu, z_xl = [z + ad * dz for z, dz in zip(Z, dZ)]
            if o.verbose: print(f"      alpha_p={a:.3f} alpha_d={ad:.3f} mu_t={mu_t:.2e}")
            delta = max(o.delta_min, delta * 0.3)
            if hold > 0:
                hold -= 1
            elif rho_p > o.rho0:
                rho_p = max(o.rho0, rho_p / o.rho_dec, o.rho_floor_frac * rho_need)
            continue
        # predictor (affine): rc = -s z
        aff = solve_dir(*[-s * z for s, z in zip(S, Z)])
        a_aff = min(max_step(S, aff[2], MK), max_step(Z, aff[1], MK))
        mu_aff = sum(((s + a_aff * ds) * mk) @ ((z + a_aff * dz) * mk) for s, ds, z, dz, mk in zip(S, aff[2], Z, aff[1], MK)) / max(nin, 1)
        sigma = (mu_aff / mu) ** 3 if mu > 0 else 0.0
        # corrector
        rc = [sigma * mu - s * z - ds * dz for s, z, ds, dz in zip(S, Z, aff[2], aff[1])]
        dx, dZ, dS, dy, dyx = solve_dir(*rc)
        a = min(1.0, o.tau * min(max_step(S, dS, MK), max_step(Z, dZ, MK)))
        x = x + a * dx
        y = y + a * dy
        yx = yx + a * dyx
        s_ru, s_rl, s_xu, s_xl = [s + a * ds for s, ds in zip(S, dS)]
        z_ru, z_rl, z_xu, z_xl = [z + a * dz for z, dz in zip(Z, dZ)]
        delta = max(o.delta_min, min(delta, 0.1 * mu / c if False else delta * 0.3))
        if rho_p > o.rho0:
            rho_p = max(o.rho0, rho_p / 3.0)
    lam_row = np.where(eqr, y, z_ru - z_rl)
    lam_box = np.where(eqx, yx, z_xu - z_xl)
    return {"status": status, "x": D * x, "yc": E * lam_row / c, "yb": lam_box / (D * c),
            "info": dict(iters=it, nfact=nfact, rp=rp, rd=rd, mu=mu, rho_p=rho_p)}
