"""DEVELOPMENT TOOL ONLY -- numpy prototype of the interior-point WARM START on "only the box changed" re-solves
(sqp_trust_region.jl:134, 574-577: after a rejected step the same QP is solved again with a smaller Delta).

Same algorithm as tests/devtools/proto_ipm.py (monotone rule, growth 4 / decay 3 inertia correction); adds
  * `state` in the result (scaled iterate: x, slacks, duals, Ruiz scaling) and
  * `warm=state`: start from the previous iterate pushed into the new box -- x clipped `kappa` (relative to the
    box width) inside, every slack >= kappa_s, every dual >= kappa_z, barrier parameter reset to `mu_w`.
Usage: python tests/devtools/proto_warm.py /tmp/pairs118.pkl
"""
from __future__ import annotations

import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, "/root/repo")
sys.path.insert(0, "/root/repo/tests/devtools")
from proto_admm import ruiz, _ninf  # noqa: E402
from proto_ipm import chol_solve  # noqa: E402


def ipm(P, q, J, rl, ru, xl, xu, warm=None, kappa=1e-2, kappa_s=1e-2, kappa_z=1e-2, mu_w=1e-3, eps=1e-9, max_iter=120,
        delta0=1e-6, delta_min=1e-8, rho0=1e-8, tau_min=0.995, verbose=False):
    n, m = q.shape[0], J.shape[0]
    P = sp.csr_matrix(P); J = sp.csr_matrix(J)
    if warm is None:
        D, E, c = ruiz(P, J, q, 15)
    else:
        D, E, c = warm["D"], warm["E"], warm["c"]
    Ps = (sp.diags(D) @ P @ sp.diags(D) * c).tocsr()
    Js = (sp.diags(E) @ J @ sp.diags(D)).tocsr(); JsT = Js.T.tocsr()
    qs = c * D * q
    rls, rus, xls, xus = E * rl, E * ru, xl / D, xu / D
    eqr, eqx = rls == rus, xls == xus
    ru_f, rl_f = np.isfinite(rus) & ~eqr, np.isfinite(rls) & ~eqr
    xu_f, xl_f = np.isfinite(xus) & ~eqx, np.isfinite(xls) & ~eqx
    MK = (ru_f, rl_f, xu_f, xl_f)
    nin = int(sum(mk.sum() for mk in MK))
    if warm is None:
        x = np.clip(np.zeros(n), xls, xus)
        Ax = Js @ x
        S = [np.where(ru_f, np.maximum(rus - Ax, 1.0), 1.0), np.where(rl_f, np.maximum(Ax - rls, 1.0), 1.0),
             np.where(xu_f, np.maximum(xus - x, 1.0), 1.0), np.where(xl_f, np.maximum(x - xls, 1.0), 1.0)]
        Z = [np.where(mk, 1.0, 0.0) for mk in MK]
        y, yx = np.zeros(m), np.zeros(n)
        mu_t = 1.0
        rho_p, rho_last = rho0, 0.0
    else:
        w = xus - xls
        marg = np.where(np.isfinite(w), np.minimum(kappa, 0.25 * w), kappa)
        x = np.minimum(np.maximum(warm["x"], np.where(np.isfinite(xls), xls + marg, -np.inf)), np.where(np.isfinite(xus), xus - marg, np.inf))
        x = np.where(eqx, xls, x)
        Ax = Js @ x
        gaps = (rus - Ax, Ax - rls, xus - x, x - xls)
        S = [np.where(mk, np.maximum(g, kappa_s), 1.0) for g, mk in zip(gaps, MK)]
        Z = [np.where(mk, np.maximum(z, kappa_z), 0.0) for z, mk in zip(warm["Z"], MK)]
        # a side that was a free inequality before and is an equality now (or vice versa) simply restarts
        y, yx = np.where(eqr, warm["y"], 0.0), np.where(eqx, warm["yx"], 0.0)
        mu_t = mu_w
        rho_p, rho_last = max(rho0, warm["rho_p"]), warm["rho_last"]
    delta = delta0
    nfact = 0
    status = "MAX_ITER"
    for it in range(max_iter):
        Ax = Js @ x; Px = Ps @ x
        lam_row = np.where(eqr, y, Z[0] - Z[1]); lam_box = np.where(eqx, yx, Z[2] - Z[3])
        r_x = Px + qs + JsT @ lam_row + lam_box
        r_eq = np.where(eqr, Ax - rls, 0.0); r_eqx = np.where(eqx, x - xls, 0.0)
        R = [np.where(ru_f, Ax + S[0] - rus, 0.0), np.where(rl_f, -Ax + S[1] + rls, 0.0), np.where(xu_f, x + S[2] - xus, 0.0),
             np.where(xl_f, -x + S[3] + xls, 0.0)]
        rp = max(_ninf(r_eq / E), _ninf(R[0] / E), _ninf(R[1] / E), _ninf(r_eqx * D), _ninf(R[2] * D), _ninf(R[3] * D))
        rd = _ninf(r_x / D) / c
        scale_p = max(1.0, _ninf(Ax / E), _ninf(x * D))
        scale_d = max(1.0, _ninf(Px / D) / c, _ninf(qs / D) / c, _ninf((JsT @ lam_row) / D) / c)
        compmax = max(_ninf((s * z) * mk) for s, z, mk in zip(S, Z, MK)) / c
        ymx = max(_ninf(lam_row), 0.0)
        if verbose:
            print(f"  it {it:3d} rp={rp:.2e} rd={rd:.2e} comp={compmax:.2e} mu_t={mu_t:.1e} rho={rho_p:.1e}")
        if rp <= eps * scale_p and rd <= eps * scale_d and compmax <= eps * max(1.0, ymx / c / 100.0):
            status = "SOLVED"
            break
        while True:
            comp = max(_ninf((s * z - mu_t) * mk) for s, z, mk in zip(S, Z, MK))
            e_mu = max(_ninf(r_x), _ninf(r_eq), _ninf(r_eqx), max(_ninf(r) for r in R), comp)
            if e_mu <= 10.0 * mu_t and mu_t > 1e-14:
                mu_t = max(1e-14, min(0.2 * mu_t, mu_t ** 1.5))
            else:
                break
        Dn = [s + delta * z for s, z in zip(S, Z)]
        w_row = np.where(eqr, 1.0 / delta, np.where(ru_f, Z[0] / Dn[0], 0.0) + np.where(rl_f, Z[1] / Dn[1], 0.0))
        w_box = np.where(eqx, 1.0 / delta, np.where(xu_f, Z[2] / Dn[2], 0.0) + np.where(xl_f, Z[3] / Dn[3], 0.0))
        while True:
            K = (Ps + sp.diags(rho_p + w_box) + JsT @ sp.diags(w_row) @ Js).tocsc()
            lu, ok = chol_solve(K, None)
            nfact += 1
            if ok:
                break
            rho_p = max(4.0 * rho_p, 1e-4 if rho_last == 0.0 else rho_last / 3.0, 1e-6)
            if rho_p > 1e8:
                return {"status": "NUMERICAL", "iters": it, "nfact": nfact}
        if rho_p > rho0 * 10:
            rho_last = rho_p
        rc = [mu_t - s * z for s, z in zip(S, Z)]
        sg = (1.0, -1.0, 1.0, -1.0)
        t_row = np.where(eqr, r_eq / delta, np.where(ru_f, (rc[0] + Z[0] * R[0]) / Dn[0], 0.0) - np.where(rl_f, (rc[1] + Z[1] * R[1]) / Dn[1], 0.0))
        t_box = np.where(eqx, r_eqx / delta, np.where(xu_f, (rc[2] + Z[2] * R[2]) / Dn[2], 0.0) - np.where(xl_f, (rc[3] + Z[3] * R[3]) / Dn[3], 0.0))
        rhs = -r_x - JsT @ t_row - t_box
        dx = lu.solve(rhs)
        Jdx = Js @ dx
        G = (Jdx, -Jdx, dx, -dx)
        dZ = [np.where(mk, (rc_ + z * (r + g)) / d, 0.0) for rc_, z, r, g, d, mk in zip(rc, Z, R, G, Dn, MK)]
        dS = [np.where(mk, -r - g + delta * dz, 0.0) for r, g, dz, mk in zip(R, G, dZ, MK)]
        dy = np.where(eqr, (Jdx + r_eq) / delta, 0.0); dyx = np.where(eqx, (dx + r_eqx) / delta, 0.0)
        a = 1.0
        for v, dv, mk in list(zip(S, dS, MK)) + list(zip(Z, dZ, MK)):
            neg = mk & (dv < 0)
            if neg.any():
                a = min(a, float(np.min(-v[neg] / dv[neg])))
        a = min(1.0, max(tau_min, 1.0 - mu_t) * a)
        x = x + a * dx; y = y + a * dy; yx = yx + a * dyx
        S = [s + a * ds for s, ds in zip(S, dS)]; Z = [z + a * dz for z, dz in zip(Z, dZ)]
        delta = max(delta_min, delta * 0.3)
        if rho_p > rho0:
            rho_p = max(rho0, rho_p / 3.0)
    lam_row = np.where(eqr, y, Z[0] - Z[1]); lam_box = np.where(eqx, yx, Z[2] - Z[3])
    return {"status": status, "iters": it, "nfact": nfact, "x": D * x, "obj": 0.5 * (D * x) @ (P @ (D * x)) + q @ (D * x),
            "yc": E * lam_row / c, "yb": lam_box / (D * c),
            "state": dict(x=x, S=S, Z=Z, y=y, yx=yx, D=D, E=E, c=c, rho_p=rho_p, rho_last=rho_last)}


if __name__ == "__main__":
    import pickle
    from oracle.coo import CooMatrix, SymCooMatrix
    from oracle.subproblem import trust_region_box
    from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
    from sqpsolver_jl_b200.nlp.networks import synth_net
    pairs = pickle.load(open(sys.argv[1], "rb"))
    net = synth_net(118, 186, 54, 118)
    nlp0 = AcopfPolar(net)

    def qp(t):
        nlp = AcopfPolar(net, pd=t["pd"], qd=t["qd"]) if "pd" in t else nlp0  # pairs of the perturbed-load batch carry their loads
        J = CooMatrix(nlp.j_row, nlp.j_col, nlp.m, nlp.n); J.fill(t["dE"])
        H = SymCooMatrix(nlp.h_row, nlp.h_col, nlp.n); H.fill(t["h_val"])
        lb, ub = trust_region_box(nlp.x_L - t["x"], nlp.x_U - t["x"], t["Delta"])
        return H.to_scipy(), t["df"], J.to_scipy(), nlp.g_L - t["E"], nlp.g_U - t["E"], lb, ub

    variants = [dict(kappa=1e-2, kappa_s=1e-2, kappa_z=1e-2, mu_w=1e-4), dict(kappa=1e-3, kappa_s=1e-3, kappa_z=1e-3, mu_w=1e-6),
                dict(kappa=1e-2, kappa_s=1e-2, kappa_z=1e-3, mu_w=1e-5), dict(kappa=1e-1, kappa_s=1e-1, kappa_z=1e-1, mu_w=1e-2),
                dict(kappa=3e-2, kappa_s=3e-2, kappa_z=3e-2, mu_w=1e-3)]
    tot = {"cold": [0, 0]}
    for a, b in pairs[: int(sys.argv[2]) if len(sys.argv) > 2 else 100]:
        r0 = ipm(*qp(a))
        cold = ipm(*qp(b))
        tot["cold"][0] += cold["iters"]; tot["cold"][1] += cold["nfact"]
        line = f"pair it{a['iter']:3d} D {a['Delta']:.2e}->{b['Delta']:.2e} first {r0['iters']}/{r0['nfact']} cold {cold['status']} {cold['iters']}/{cold['nfact']}"
        for v in variants:
            w = ipm(*qp(b), warm=r0["state"], **v)
            key = str(v)
            tot.setdefault(key, [0, 0, 0.0])
            tot[key][0] += w["iters"]; tot[key][1] += w["nfact"]
            dobj = (w.get("obj", np.nan) - cold["obj"]) / max(1.0, abs(cold["obj"]))
            tot[key][2] = max(tot[key][2], abs(dobj) if np.isfinite(dobj) else 9.9)
            line += f" | {w['status'][:3]} {w['iters']}/{w['nfact']} dobj {dobj:+.1e}"
        print(line, flush=True)
    for k, v in tot.items():
        print(k, v)
