// ipm.cuh -- regularised primal-dual interior point method on the scaled QP, one team per instance.
//
// Why it exists next to ADMM: the SQP subproblems of ACOPF are LP-like,
// degenerate and -- after the first iteration -- nonconvex; ADMM needs thousands of
// iterations on them and its tail is too slow to reach the 1e-8 steps the SQP termination
// test needs.  The interior point method solves them in 10-60 Newton steps to 1e-9, and it is
// the algorithm family of the sub-solver the reference uses (Ipopt), so it picks the same
// kind of solution (analytic-centre multipliers).  Its Newton system is CONDENSED to
//
//      K dx = rhs,   K = P + diag(rho_p + w_box) + J' diag(w_row) J      (always SPD after
//                                                                         inertia correction)
// which has the structure of the ADMM matrix, so the scaled CSR values, the SpMV row kernels
// and the reductions of admm.cuh/team.cuh are reused; K is factorised by chol.cuh.
//
//   inequality sides k (finite row/box bounds):  g_k' x + s_k = h_k,  s_k, z_k > 0
//   equality rows / fixed columns: dual-regularised,  A dx - delta dy = -r
//   weights: w = z / (s + delta z) per finite side, 1/delta on equalities
//   Mehrotra predictor-corrector, fraction-to-the-boundary step, delta -> delta_min geometrically
//   indefinite P: Cholesky breaks down -> rho_p is raised until it succeeds (Ipopt's inertia rule)
#pragma once
#include "admm.cuh"
#include "chol.cuh"


struct SideDir { double dz, ds; };
__device__ __forceinline__ SideDir side_dir(double rc, double z, double s, double r, double gdx, double delta) {
    double d = s + delta * z;
    SideDir o;
    o.dz = (rc + z * (r + gdx)) / d;
    o.ds = -r - gdx + delta * o.dz;
    return o;
}
// largest step fraction bookkeeping: returns max(-dv/v, 0) so that alpha = 1/max(...)
__device__ __forceinline__ double step_ratio(double v, double dv) { return dv < 0.0 ? -dv / v : 0.0; }

template <class Team>
__device__ IpmOut ipm_run(Team& T, const Inst& I, const CholDev& C, double* Lval, double* yw, const sqpqp_options& o,
                          double c, int phase, const double* xk_scaled_start) {
    const int N = I.N, M = I.M;
    // scaled problem data (set up by solve_instance)
    const double *q = I.nv[N_Q], *xl = I.nv[N_XL], *xu = I.nv[N_XU], *D = I.nv[N_D], *hd = I.nv[N_HD];
    const double *rl = I.mv[M_RL], *ru = I.mv[M_RU], *Es = I.mv[M_ES];
    // IPM state (aliases of the ADMM workspace slots; the two methods never run concurrently)
    double *x = I.nv[N_X], *dx = I.nv[N_XT], *rx = I.nv[N_R], *rhs = I.nv[N_P];
    double *sxu = I.nv[N_ZB], *zxu = I.nv[N_YB], *sxl = I.nv[N_RB], *zxl = I.nv[N_KP];
    double *cxu = I.nv[N_MINV], *cxl = I.nv[N_XFIX], *yx = I.nv[N_MASK], *wb = I.nv[N_DSH], *tmpN = I.nv[N_TMP];
    double *sru = I.mv[M_ZC], *zru = I.mv[M_YC], *srl = I.mv[M_RC], *zrl = I.mv[M_BC];
    double *cru = I.mv[M_YP], *crl = I.mv[M_TMP], *y = I.mv[M_I1], *w = I.mv[M_RW], *t = I.mv[M_T];
    double *jdx = I.mv[M_I2], *Ax = I.mv[M_I3];
    // side residuals are TRACKED (r += alpha (g dx + ds)), never recomputed from A x - b: at the end
    // of the solve delta ~ 1e-11 and a freshly evaluated residual carries ~1e-16 |Ax| of rounding
    // noise, which the dual update dy = (J dx + r)/delta would amplify by 1e11
    double *rru = I.mv[M_AX], *rrl = I.mv[M_I4], *rxu = I.nv[N_TMP2], *rxl = I.nv[N_I1];

    IpmOut out{false, false, false, false, 0, 0, INFINITY, INFINITY, 0.0};
    // ---- start point: x inside the box, unit duals, slacks >= 1 --------------------------------
    for_n(T, N, [&](int j) {
        double v = xk_scaled_start ? xk_scaled_start[j] : 0.0;
        v = fmin(fmax(v, xl[j]), xu[j]);
        x[j] = v;
        yx[j] = 0.0;
    });
    T.sync();
    double cnt[1] = {0.0};
    csr_rows(T, M, I.lgJ, I.J.rb, I.J.re, I.J.col, I.Jsv, x, [&](int i, double ax) {
        bool eq = rl[i] == ru[i];
        bool uf = !eq && !isinf(ru[i]), lf = !eq && !isinf(rl[i]);
        sru[i] = uf ? fmax(ru[i] - ax, 1.0) : 1.0; zru[i] = uf ? 1.0 : 0.0;
        srl[i] = lf ? fmax(ax - rl[i], 1.0) : 1.0; zrl[i] = lf ? 1.0 : 0.0;
        y[i] = 0.0;
        Ax[i] = ax;
        rru[i] = eq ? ax - rl[i] : (uf ? ax + sru[i] - ru[i] : 0.0);
        rrl[i] = lf ? -ax + srl[i] + rl[i] : 0.0;
        cnt[0] += (double)uf + (double)lf;
    });
    for_n(T, N, [&](int j) {
        bool eq = xl[j] == xu[j];
        bool uf = !eq && !isinf(xu[j]), lf = !eq && !isinf(xl[j]);
        sxu[j] = uf ? fmax(xu[j] - x[j], 1.0) : 1.0; zxu[j] = uf ? 1.0 : 0.0;
        sxl[j] = lf ? fmax(x[j] - xl[j], 1.0) : 1.0; zxl[j] = lf ? 1.0 : 0.0;
        rxu[j] = eq ? x[j] - xl[j] : (uf ? x[j] + sxu[j] - xu[j] : 0.0);
        rxl[j] = lf ? -x[j] + sxl[j] + xl[j] : 0.0;
        cnt[0] += (double)uf + (double)lf;
    });
    T.template reduce<1, false>(cnt);
    const double nin = fmax(cnt[0], 1.0);

    double delta = o.ipm_delta0, rho_p = o.ipm_rho0, rho_last = 0.0;
    int acc_cnt = 0;
    double rp_ref = INFINITY;
    double mu_t = o.ipm_mu0;  // barrier parameter of the current subproblem (monotone Fiacco-McCormick rule)
    for (int it = 0; it < o.ipm_max_iter; ++it) {
        out.iters = it;
        // ---- residuals ----------------------------------------------------------------------
        T.sync();
        double ymx[1] = {0.0};
        for_n(T, M, [&](int i) { double v = (rl[i] == ru[i]) ? y[i] : (zru[i] - zrl[i]); t[i] = v; ymx[0] = fmax(ymx[0], fabs(v)); });
        T.template reduce<1, true>(ymx);
        if (ymx[0] > 1e12) { out.blowup = true; break; }  // multipliers exploding: ADMM certifies infeasibility
        double mx[6] = {0, 0, 0, 0, 0, 0};  // [0] rp [1] rd*c [2] primal scale [3] dual scale*c [4] |A'lam|*c [5] |lam|*c (unscaled)
        double sup[1] = {0.0};  // support function  u'(lam)+ + l'(lam)-  (scaled units = unscaled * c)
        double sm[1] = {0.0};         // sum s*z
        double sz[3] = {0.0, -INFINITY, 0.0};  // [0] max s*z  [1] max -(s*z) = -min s*z  [2] scaled residual of the barrier problem
        csr_rows2(T, N, I.lgT, I.H.rb, I.H.re, I.H.col, I.Hsv, x, I.useH, I.T.rb, I.T.re, I.T.col, I.Tsv, t,
                  [&](int j, double px, double aty) {
                      if (!I.useH) px = hd[j] * x[j];
                      bool eq = xl[j] == xu[j];
                      double lamb = eq ? yx[j] : (zxu[j] - zxl[j]);
                      double r = px + q[j] + aty + lamb;
                      rx[j] = r;
                      mx[4] = fmax(mx[4], fabs(aty + lamb) / D[j]);
                      mx[5] = fmax(mx[5], fabs(lamb) / D[j]);
                      if (lamb > 0.0) sup[0] += xu[j] * lamb; else if (lamb < 0.0) sup[0] += xl[j] * lamb;
                      sz[2] = fmax(sz[2], fabs(r));
                      double id = 1.0 / D[j];
                      mx[1] = fmax(mx[1], fabs(r) * id);
                      if (o.verbose > 1 && it >= 9 && fabs(r) * id / c > 1e-4)
                          printf("      j=%d r=%.3e px=%.3e q=%.3e aty=%.3e lamb=%.3e x=%.6e xl=%.6e xu=%.6e sxu=%.3e zxu=%.3e sxl=%.3e zxl=%.3e D=%.2e\n",
                                 j, r, px, q[j], aty, lamb, x[j], xl[j], xu[j], sxu[j], zxu[j], sxl[j], zxl[j], D[j]);
                      mx[3] = fmax(mx[3], fmax(fabs(px), fmax(fabs(aty), fabs(q[j]))) * id);
                      mx[2] = fmax(mx[2], fabs(x[j]) * D[j]);
                      double pr = 0.0;
                      if (eq) pr = fabs(rxu[j]);
                      else {
                          if (!isinf(xu[j])) { double p_ = sxu[j] * zxu[j]; pr = fmax(pr, fabs(rxu[j])); sm[0] += p_; sz[0] = fmax(sz[0], p_); sz[1] = fmax(sz[1], -p_); }
                          if (!isinf(xl[j])) { double p_ = sxl[j] * zxl[j]; pr = fmax(pr, fabs(rxl[j])); sm[0] += p_; sz[0] = fmax(sz[0], p_); sz[1] = fmax(sz[1], -p_); }
                      }
                      sz[2] = fmax(sz[2], pr);
                      mx[0] = fmax(mx[0], pr * D[j]);
                  });
        for_n(T, M, [&](int i) {
            bool eq = rl[i] == ru[i];
            double pr = 0.0, ax = Ax[i];
            if (eq) pr = fabs(rru[i]);
            else {
                if (!isinf(ru[i])) { double p_ = sru[i] * zru[i]; pr = fmax(pr, fabs(rru[i])); sm[0] += p_; sz[0] = fmax(sz[0], p_); sz[1] = fmax(sz[1], -p_); }
                if (!isinf(rl[i])) { double p_ = srl[i] * zrl[i]; pr = fmax(pr, fabs(rrl[i])); sm[0] += p_; sz[0] = fmax(sz[0], p_); sz[1] = fmax(sz[1], -p_); }
            }
            sz[2] = fmax(sz[2], pr);
            mx[0] = fmax(mx[0], pr / Es[i]);
            mx[2] = fmax(mx[2], fabs(ax) / Es[i]);
            double lr = t[i];
            mx[5] = fmax(mx[5], fabs(lr) * Es[i]);
            if (lr > 0.0) sup[0] += ru[i] * lr; else if (lr < 0.0) sup[0] += rl[i] * lr;
        });
        T.template reduce<6, true>(mx);
        T.template reduce<1, false>(sup);
        T.template reduce<1, false>(sm);
        // -sz[1] = min s*z: initialise properly (max of negatives starts at -inf)
        T.template reduce<3, true>(sz);
        const double mu = sm[0] / nin;
        out.rp = mx[0];
        out.rd = mx[1] / c;
        const double scale_p = fmax(1.0, mx[2]), scale_d = fmax(1.0, mx[3] / c);
        // primal infeasibility certificate on the multiplier direction (same test as the ADMM path applies
        // to its dual increments): A' lam ~ 0 while the support function of the bounds is negative
        if (mx[5] / c > 1e4 && mx[4] <= o.eps_inf * mx[5] && sup[0] <= -o.eps_inf * mx[5]) {
            out.infeasible = true;
            break;
        }
        // ... or, for marginally infeasible rows (violation << 1, multipliers growing only linearly):
        // the regularised equalities make the iterates converge to a minimiser of the violation, so a
        // primal residual that has stalled at a positive value while the dual residual is converged is
        // the interior-point analogue of Ipopt's failed restoration phase -> LOCALLY_INFEASIBLE
        if (it >= 20 && it % 10 == 0) {
            bool stalled = out.rp > 1e4 * o.ipm_eps * scale_p && fabs(out.rp - rp_ref) <= 1e-3 * out.rp &&
                           out.rd <= 1e-5 * scale_d && mx[5] / c > 1e3;
            if (stalled) { out.infeasible = true; break; }
        }
        if (it % 10 == 0) rp_ref = out.rp;
        // Termination (Ipopt-style scaling): primal residual relative to |x|,|Ax|; stationarity relative to
        // the gradient terms (its attainable floor is ~1e-9 of them, the conditioning of K); complementarity
        // (largest s*z, unscaled) absolute unless the multipliers themselves are large.
        const double comp_u = sz[0] / c, sc = fmax(1.0, ymx[0] / c / 100.0);
        if (out.rp <= o.ipm_eps * scale_p && out.rd <= o.ipm_eps * scale_d && comp_u <= o.ipm_eps * sc) {
            out.solved = true;
            break;
        }
        const double acc_eps = 100.0 * o.ipm_eps;
        bool acceptable = out.rp <= acc_eps * scale_p && out.rd <= acc_eps * scale_d && comp_u <= acc_eps * sc;
        acc_cnt = acceptable ? acc_cnt + 1 : 0;
        out.almost = out.rp <= 1e-6 * scale_p && out.rd <= 1e-6 * scale_d && comp_u <= 1e-6 * sc;
        if (acc_cnt >= 8) {  // stuck on the floor of an acceptable point
            out.solved = true;
            break;
        }
        if (o.verbose) {
            double ym[3] = {0, 0, 0};
            for_n(T, M, [&](int i) { ym[0] = fmax(ym[0], fabs(y[i])); ym[1] = fmax(ym[1], fmax(zru[i], zrl[i])); });
            for_n(T, N, [&](int j) { ym[1] = fmax(ym[1], fmax(zxu[j], zxl[j])); ym[2] = fmax(ym[2], fabs(x[j])); });
            T.template reduce<3, true>(ym);
            if (T.tid() == 0) printf("      |y|=%.2e |z|=%.2e |x|=%.2e c=%.2e\n", ym[0], ym[1], ym[2], c);
        }
        if (o.verbose && T.tid() == 0)
            printf("  ipm %3d rp=%.2e rd=%.2e mu=%.2e delta=%.1e rho=%.1e nfact=%d\n", it, out.rp, out.rd, mu / c, delta, rho_p, out.nfact);
        if (!(mu == mu) || !(out.rd == out.rd)) break;  // NaN guard
        // ---- barrier update: shrink mu_t while the current barrier problem is solved to kappa*mu_t ---
        for (int g = 0; g < 60; ++g) {
            double comp = fmax(fabs(sz[0] - mu_t), fabs(-sz[1] - mu_t));
            double e_mu = fmax(sz[2], comp);
            if (e_mu <= o.ipm_kappa_eps * mu_t && mu_t > o.ipm_mu_min) mu_t = fmax(o.ipm_mu_min, fmin(0.2 * mu_t, mu_t * sqrt(mu_t)));
            else break;
        }
        // ---- weights, assembly, factorisation (with inertia correction) ---------------------------
        for_n(T, M, [&](int i) {
            bool eq = rl[i] == ru[i];
            double wi = 0.0;
            if (eq) wi = 1.0 / delta;
            else {
                if (!isinf(ru[i])) wi += zru[i] / (sru[i] + delta * zru[i]);
                if (!isinf(rl[i])) wi += zrl[i] / (srl[i] + delta * zrl[i]);
            }
            w[i] = wi;
        });
        bool fact_ok = false;
        for (int tries = 0; tries < 30 && !fact_ok; ++tries) {
            for_n(T, N, [&](int j) {
                bool eq = xl[j] == xu[j];
                double wj = 0.0;
                if (eq) wj = 1.0 / delta;
                else {
                    if (!isinf(xu[j])) wj += zxu[j] / (sxu[j] + delta * zxu[j]);
                    if (!isinf(xl[j])) wj += zxl[j] / (sxl[j] + delta * zxl[j]);
                }
                wb[j] = wj + rho_p + (I.useH ? 0.0 : hd[j]);
            });
            T.sync();
            chol_assemble(T, C, Lval, I.useH ? I.Hsv : (const double*)nullptr, wb, w, I.Jsv);
            fact_ok = chol_factor(T, C, Lval);
            ++out.nfact;
            if (!fact_ok) rho_p = fmax(fmax(10.0 * rho_p, rho_last > 0.0 ? rho_last / 3.0 : 1e-4), 1e-6);
            if (rho_p > 1e8) break;
        }
        if (!fact_ok) break;
        if (rho_p > 10.0 * o.ipm_rho0) rho_last = rho_p;
        out.rho_p = rho_p;

        // ---- two Newton solves: predictor (sigma = 0, no cross term), corrector ----------------------
        double sigma_mu = mu_t, alpha = 1.0;
        const double tau_k = fmax(o.ipm_tau, 1.0 - mu_t);
        for (int pass = 1; pass < 2; ++pass) {
            // t_row, then rhs = -r_x - T t_row - t_box
            for_n(T, M, [&](int i) {
                bool eq = rl[i] == ru[i];
                double ax = Ax[i], ti = 0.0;
                if (eq) ti = rru[i] / delta;
                else {
                    if (!isinf(ru[i])) {
                        double rc = sigma_mu - sru[i] * zru[i] - 0.0;
                        ti += (rc + zru[i] * (rru[i])) / (sru[i] + delta * zru[i]);
                    }
                    if (!isinf(rl[i])) {
                        double rc = sigma_mu - srl[i] * zrl[i] - 0.0;
                        ti -= (rc + zrl[i] * (rrl[i])) / (srl[i] + delta * zrl[i]);
                    }
                }
                t[i] = ti;
            });
            T.sync();
            csr_rows(T, N, I.lgT, I.T.rb, I.T.re, I.T.col, I.Tsv, t, [&](int j, double tt) {
                bool eq = xl[j] == xu[j];
                double tb = 0.0;
                if (eq) tb = rxu[j] / delta;
                else {
                    if (!isinf(xu[j])) {
                        double rc = sigma_mu - sxu[j] * zxu[j] - 0.0;
                        tb += (rc + zxu[j] * (rxu[j])) / (sxu[j] + delta * zxu[j]);
                    }
                    if (!isinf(xl[j])) {
                        double rc = sigma_mu - sxl[j] * zxl[j] - 0.0;
                        tb -= (rc + zxl[j] * (rxl[j])) / (sxl[j] + delta * zxl[j]);
                    }
                }
                rhs[j] = -rx[j] - tt - tb;
            });
            T.sync();
            chol_solve(T, C, Lval, rhs, dx, yw);
            // iterative refinement against the matrix-free K (K is ill-conditioned by design)
            for (int rf = 0; rf < o.ipm_refine; ++rf) {
                apply_K(T, I, dx, tmpN, wb, w, (const double*)nullptr);
                T.sync();
                // apply_K adds hd for !useH on top of dsh; wb already contains it -> subtract once
                double nr[2] = {0.0, 0.0};
                for_n(T, N, [&](int j) {
                    double kv = tmpN[j] - (I.useH ? 0.0 : hd[j] * dx[j]);
                    double r = rhs[j] - kv;
                    tmpN[j] = r;
                    nr[0] = fmax(nr[0], fabs(r));
                    nr[1] = fmax(nr[1], fabs(rhs[j]));
                });
                if (o.verbose) {
                    T.template reduce<2, true>(nr);
                    if (T.tid() == 0) printf("      pass %d refine %d: |rhs - K dx| = %.2e  |rhs| = %.2e\n", pass, rf, nr[0], nr[1]);
                }
                T.sync();
                chol_solve(T, C, Lval, tmpN, tmpN, yw);
                for_n(T, N, [&](int j) { dx[j] += tmpN[j]; });
                T.sync();
            }
            // J dx and the step-to-boundary ratio
            double ratio[1] = {0.0};
            csr_rows(T, M, I.lgJ, I.J.rb, I.J.re, I.J.col, I.Jsv, dx, [&](int i, double jd) {
                jdx[i] = jd;
                if (rl[i] == ru[i]) return;
                double ax = Ax[i];
                if (!isinf(ru[i])) {
                    double rc = sigma_mu - sru[i] * zru[i] - 0.0;
                    SideDir d = side_dir(rc, zru[i], sru[i], rru[i], jd, delta);
                    ratio[0] = fmax(ratio[0], fmax(step_ratio(sru[i], d.ds), step_ratio(zru[i], d.dz)));
                }
                if (!isinf(rl[i])) {
                    double rc = sigma_mu - srl[i] * zrl[i] - 0.0;
                    SideDir d = side_dir(rc, zrl[i], srl[i], rrl[i], -jd, delta);
                    ratio[0] = fmax(ratio[0], fmax(step_ratio(srl[i], d.ds), step_ratio(zrl[i], d.dz)));
                }
            });
            for_n(T, N, [&](int j) {
                if (xl[j] == xu[j]) return;
                if (!isinf(xu[j])) {
                    double rc = sigma_mu - sxu[j] * zxu[j] - 0.0;
                    SideDir d = side_dir(rc, zxu[j], sxu[j], rxu[j], dx[j], delta);
                    ratio[0] = fmax(ratio[0], fmax(step_ratio(sxu[j], d.ds), step_ratio(zxu[j], d.dz)));
                }
                if (!isinf(xl[j])) {
                    double rc = sigma_mu - sxl[j] * zxl[j] - 0.0;
                    SideDir d = side_dir(rc, zxl[j], sxl[j], rxl[j], -dx[j], delta);
                    ratio[0] = fmax(ratio[0], fmax(step_ratio(sxl[j], d.ds), step_ratio(zxl[j], d.dz)));
                }
            });
            T.template reduce<1, true>(ratio);
            if (pass == 0) {
                double a_aff = ratio[0] > 1.0 ? 1.0 / ratio[0] : 1.0;
                // mu_aff and the second-order cross terms ds_aff * dz_aff
                double ms[1] = {0.0};
                for_n(T, M, [&](int i) {
                    if (rl[i] == ru[i]) return;
                    double ax = Ax[i], jd = jdx[i];
                    if (!isinf(ru[i])) {
                        SideDir d = side_dir(-sru[i] * zru[i], zru[i], sru[i], rru[i], jd, delta);
                        cru[i] = d.ds * d.dz;
                        ms[0] += (sru[i] + a_aff * d.ds) * (zru[i] + a_aff * d.dz);
                    }
                    if (!isinf(rl[i])) {
                        SideDir d = side_dir(-srl[i] * zrl[i], zrl[i], srl[i], rrl[i], -jd, delta);
                        crl[i] = d.ds * d.dz;
                        ms[0] += (srl[i] + a_aff * d.ds) * (zrl[i] + a_aff * d.dz);
                    }
                });
                for_n(T, N, [&](int j) {
                    if (xl[j] == xu[j]) return;
                    if (!isinf(xu[j])) {
                        SideDir d = side_dir(-sxu[j] * zxu[j], zxu[j], sxu[j], rxu[j], dx[j], delta);
                        cxu[j] = d.ds * d.dz;
                        ms[0] += (sxu[j] + a_aff * d.ds) * (zxu[j] + a_aff * d.dz);
                    }
                    if (!isinf(xl[j])) {
                        SideDir d = side_dir(-sxl[j] * zxl[j], zxl[j], sxl[j], rxl[j], -dx[j], delta);
                        cxl[j] = d.ds * d.dz;
                        ms[0] += (sxl[j] + a_aff * d.ds) * (zxl[j] + a_aff * d.dz);
                    }
                });
                T.template reduce<1, false>(ms);
                double mu_aff = ms[0] / nin;
                double sg = mu > 0.0 ? mu_aff / mu : 0.0;
                sg = fmin(fmax(sg, 0.0), 1.0);
                sigma_mu = sg * sg * sg * mu;
            } else {
                alpha = 1.0;
                if (ratio[0] > 0.0) alpha = fmin(1.0, tau_k / ratio[0]);
                if (o.verbose && T.tid() == 0) printf("      sigma_mu=%.2e alpha=%.3e\n", sigma_mu / c, alpha);
            }
        }
        // ---- update with the corrector direction -------------------------------------------------------
        for_n(T, M, [&](int i) {
            double ax = Ax[i], jd = jdx[i];
            Ax[i] = ax + alpha * jd;
            if (rl[i] == ru[i]) { y[i] += alpha * (jd + rru[i]) / delta; rru[i] += alpha * jd; return; }
            if (!isinf(ru[i])) {
                SideDir d = side_dir(sigma_mu - sru[i] * zru[i] , zru[i], sru[i], rru[i], jd, delta);
                sru[i] += alpha * d.ds; zru[i] += alpha * d.dz; rru[i] += alpha * (jd + d.ds);
            }
            if (!isinf(rl[i])) {
                SideDir d = side_dir(sigma_mu - srl[i] * zrl[i] , zrl[i], srl[i], rrl[i], -jd, delta);
                srl[i] += alpha * d.ds; zrl[i] += alpha * d.dz; rrl[i] += alpha * (-jd + d.ds);
            }
        });
        for_n(T, N, [&](int j) {
            double xj = x[j], dj = dx[j];
            if (xl[j] == xu[j]) { yx[j] += alpha * (dj + rxu[j]) / delta; rxu[j] += alpha * dj; x[j] = xj + alpha * dj; return; }
            if (!isinf(xu[j])) {
                SideDir d = side_dir(sigma_mu - sxu[j] * zxu[j] , zxu[j], sxu[j], rxu[j], dj, delta);
                sxu[j] += alpha * d.ds; zxu[j] += alpha * d.dz; rxu[j] += alpha * (dj + d.ds);
            }
            if (!isinf(xl[j])) {
                SideDir d = side_dir(sigma_mu - sxl[j] * zxl[j] , zxl[j], sxl[j], rxl[j], -dj, delta);
                sxl[j] += alpha * d.ds; zxl[j] += alpha * d.dz; rxl[j] += alpha * (-dj + d.ds);
            }
            x[j] = xj + alpha * dj;
        });
        delta = fmax(o.ipm_delta_min, delta * 0.3);
        if (rho_p > o.ipm_rho0) rho_p = fmax(o.ipm_rho0, rho_p / 3.0);
        out.iters = it + 1;
    }
    T.sync();
    if (!out.solved && acc_cnt > 0) out.solved = true;  // iteration cap reached on an acceptable point
    if (out.solved || out.almost) {
        // multipliers in the OSQP sign the output stage expects: yc (rows) and yb (box)
        double *yc = I.mv[M_YC], *yb = I.nv[N_YB];
        for_n(T, M, [&](int i) { yc[i] = (rl[i] == ru[i]) ? y[i] : (zru[i] - zrl[i]); });
        for_n(T, N, [&](int j) { yb[j] = (xl[j] == xu[j]) ? yx[j] : (zxu[j] - zxl[j]); });
        T.sync();
    }
    return out;
}
