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
__device__ IpmOut ipm_run(Team& T, const Inst& I, const CholDev& C, const CholWork& W, const sqpqp_options& o,
                          double c, int phase, const double* xk_scaled_start) {
    const int N = I.N, M = I.M;
    // Work vectors are addressed through the (shared-memory resident) pointer table of `I` at every use
    // instead of through ~36 local pointer aliases: 72 fewer live registers, which is what lets the
    // 64-register / 4-CTAs-per-SM variant of the kernel run without spilling its address arithmetic.
    // Aliases of the ADMM workspace slots (the two methods never run concurrently); side residuals are
    // TRACKED (r += alpha (g dx + ds)), never recomputed from A x - b: at the end of the solve
    // delta ~ 1e-8 and a freshly evaluated residual carries ~1e-16 |Ax| of rounding noise, which the
    // dual update dy = (J dx + r)/delta would amplify by 1/delta.
    IpmOut out{false, false, false, false, 0, 0, INFINITY, INFINITY, 0.0};
    Prof pf;
    pf.start();
    // ---- start point: I.nv[N_X] inside the box, unit duals, slacks >= 1 --------------------------------
    for_n(T, N, [&](int j) {
        double v = xk_scaled_start ? xk_scaled_start[j] : 0.0;
        v = fmin(fmax(v, I.nv[N_XL][j]), I.nv[N_XU][j]);
        I.nv[N_X][j] = v;
        I.nv[N_MASK][j] = 0.0;
    });
    T.sync();
    double cnt[1] = {0.0};
    csr_rows(T, M, I.lgJ, I.J.rb, I.J.re, I.J.col, I.Jsv, I.nv[N_X], [&](int i, double ax) {
        bool eq = I.mv[M_RL][i] == I.mv[M_RU][i];
        bool uf = !eq && !isinf(I.mv[M_RU][i]), lf = !eq && !isinf(I.mv[M_RL][i]);
        I.mv[M_ZC][i] = uf ? fmax(I.mv[M_RU][i] - ax, 1.0) : 1.0; I.mv[M_YC][i] = uf ? 1.0 : 0.0;
        I.mv[M_RC][i] = lf ? fmax(ax - I.mv[M_RL][i], 1.0) : 1.0; I.mv[M_BC][i] = lf ? 1.0 : 0.0;
        I.mv[M_I1][i] = 0.0;
        I.mv[M_I3][i] = ax;
        I.mv[M_AX][i] = eq ? ax - I.mv[M_RL][i] : (uf ? ax + I.mv[M_ZC][i] - I.mv[M_RU][i] : 0.0);
        I.mv[M_I4][i] = lf ? -ax + I.mv[M_RC][i] + I.mv[M_RL][i] : 0.0;
        cnt[0] += (double)uf + (double)lf;
    });
    for_n(T, N, [&](int j) {
        bool eq = I.nv[N_XL][j] == I.nv[N_XU][j];
        bool uf = !eq && !isinf(I.nv[N_XU][j]), lf = !eq && !isinf(I.nv[N_XL][j]);
        I.nv[N_ZB][j] = uf ? fmax(I.nv[N_XU][j] - I.nv[N_X][j], 1.0) : 1.0; I.nv[N_YB][j] = uf ? 1.0 : 0.0;
        I.nv[N_RB][j] = lf ? fmax(I.nv[N_X][j] - I.nv[N_XL][j], 1.0) : 1.0; I.nv[N_KP][j] = lf ? 1.0 : 0.0;
        I.nv[N_TMP2][j] = eq ? I.nv[N_X][j] - I.nv[N_XL][j] : (uf ? I.nv[N_X][j] + I.nv[N_ZB][j] - I.nv[N_XU][j] : 0.0);
        I.nv[N_I1][j] = lf ? -I.nv[N_X][j] + I.nv[N_RB][j] + I.nv[N_XL][j] : 0.0;
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
        for_n(T, M, [&](int i) { double v = (I.mv[M_RL][i] == I.mv[M_RU][i]) ? I.mv[M_I1][i] : (I.mv[M_YC][i] - I.mv[M_BC][i]); I.mv[M_T][i] = v; ymx[0] = fmax(ymx[0], fabs(v)); });
        T.template reduce<1, true>(ymx);
        if (ymx[0] > 1e12) { out.blowup = true; break; }  // multipliers exploding: ADMM certifies infeasibility
        double mx[6] = {0, 0, 0, 0, 0, 0};  // [0] rp [1] rd*c [2] primal scale [3] dual scale*c [4] |A'lam|*c [5] |lam|*c (unscaled)
        double sup[1] = {0.0};  // support function  u'(lam)+ + l'(lam)-  (scaled units = unscaled * c)
        double sm[1] = {0.0};         // sum s*z
        double sz[3] = {0.0, -INFINITY, 0.0};  // [0] max s*z  [1] max -(s*z) = -min s*z  [2] scaled residual of the barrier problem
        csr_rows2(T, N, I.lgT, I.H.rb, I.H.re, I.H.col, I.Hsv, I.nv[N_X], I.useH, I.T.rb, I.T.re, I.T.col, I.Tsv, I.mv[M_T],
                  [&](int j, double px, double aty) {
                      if (!I.useH) px = I.nv[N_HD][j] * I.nv[N_X][j];
                      bool eq = I.nv[N_XL][j] == I.nv[N_XU][j];
                      double lamb = eq ? I.nv[N_MASK][j] : (I.nv[N_YB][j] - I.nv[N_KP][j]);
                      double r = px + I.nv[N_Q][j] + aty + lamb;
                      I.nv[N_R][j] = r;
                      mx[4] = fmax(mx[4], fabs(aty + lamb) / I.nv[N_D][j]);
                      mx[5] = fmax(mx[5], fabs(lamb) / I.nv[N_D][j]);
                      if (lamb > 0.0) sup[0] += I.nv[N_XU][j] * lamb; else if (lamb < 0.0) sup[0] += I.nv[N_XL][j] * lamb;
                      sz[2] = fmax(sz[2], fabs(r));
                      double id = 1.0 / I.nv[N_D][j];
                      mx[1] = fmax(mx[1], fabs(r) * id);
                      if (o.verbose > 1 && it >= 9 && fabs(r) * id / c > 1e-4)
                          printf("      j=%d r=%.3e px=%.3e I.nv[N_Q]=%.3e aty=%.3e lamb=%.3e I.nv[N_X]=%.6e I.nv[N_XL]=%.6e I.nv[N_XU]=%.6e I.nv[N_ZB]=%.3e I.nv[N_YB]=%.3e I.nv[N_RB]=%.3e I.nv[N_KP]=%.3e I.nv[N_D]=%.2e\n",
                                 j, r, px, I.nv[N_Q][j], aty, lamb, I.nv[N_X][j], I.nv[N_XL][j], I.nv[N_XU][j], I.nv[N_ZB][j], I.nv[N_YB][j], I.nv[N_RB][j], I.nv[N_KP][j], I.nv[N_D][j]);
                      mx[3] = fmax(mx[3], fmax(fabs(px), fmax(fabs(aty), fabs(I.nv[N_Q][j]))) * id);
                      mx[2] = fmax(mx[2], fabs(I.nv[N_X][j]) * I.nv[N_D][j]);
                      double pr = 0.0;
                      if (eq) pr = fabs(I.nv[N_TMP2][j]);
                      else {
                          if (!isinf(I.nv[N_XU][j])) { double p_ = I.nv[N_ZB][j] * I.nv[N_YB][j]; pr = fmax(pr, fabs(I.nv[N_TMP2][j])); sm[0] += p_; sz[0] = fmax(sz[0], p_); sz[1] = fmax(sz[1], -p_); }
                          if (!isinf(I.nv[N_XL][j])) { double p_ = I.nv[N_RB][j] * I.nv[N_KP][j]; pr = fmax(pr, fabs(I.nv[N_I1][j])); sm[0] += p_; sz[0] = fmax(sz[0], p_); sz[1] = fmax(sz[1], -p_); }
                      }
                      sz[2] = fmax(sz[2], pr);
                      mx[0] = fmax(mx[0], pr * I.nv[N_D][j]);
                  });
        for_n(T, M, [&](int i) {
            bool eq = I.mv[M_RL][i] == I.mv[M_RU][i];
            double pr = 0.0, ax = I.mv[M_I3][i];
            if (eq) pr = fabs(I.mv[M_AX][i]);
            else {
                if (!isinf(I.mv[M_RU][i])) { double p_ = I.mv[M_ZC][i] * I.mv[M_YC][i]; pr = fmax(pr, fabs(I.mv[M_AX][i])); sm[0] += p_; sz[0] = fmax(sz[0], p_); sz[1] = fmax(sz[1], -p_); }
                if (!isinf(I.mv[M_RL][i])) { double p_ = I.mv[M_RC][i] * I.mv[M_BC][i]; pr = fmax(pr, fabs(I.mv[M_I4][i])); sm[0] += p_; sz[0] = fmax(sz[0], p_); sz[1] = fmax(sz[1], -p_); }
            }
            sz[2] = fmax(sz[2], pr);
            mx[0] = fmax(mx[0], pr / I.mv[M_ES][i]);
            mx[2] = fmax(mx[2], fabs(ax) / I.mv[M_ES][i]);
            double lr = I.mv[M_T][i];
            mx[5] = fmax(mx[5], fabs(lr) * I.mv[M_ES][i]);
            if (lr > 0.0) sup[0] += I.mv[M_RU][i] * lr; else if (lr < 0.0) sup[0] += I.mv[M_RL][i] * lr;
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
        // Termination (Ipopt-style scaling): primal residual relative to |I.nv[N_X]|,|I.mv[M_I3]|; stationarity relative to
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
            for_n(T, M, [&](int i) { ym[0] = fmax(ym[0], fabs(I.mv[M_I1][i])); ym[1] = fmax(ym[1], fmax(I.mv[M_YC][i], I.mv[M_BC][i])); });
            for_n(T, N, [&](int j) { ym[1] = fmax(ym[1], fmax(I.nv[N_YB][j], I.nv[N_KP][j])); ym[2] = fmax(ym[2], fabs(I.nv[N_X][j])); });
            T.template reduce<3, true>(ym);
            if (T.tid() == 0) printf("      |I.mv[M_I1]|=%.2e |z|=%.2e |I.nv[N_X]|=%.2e c=%.2e\n", ym[0], ym[1], ym[2], c);
        }
        if (o.verbose && T.tid() == 0)
            printf("  ipm %3d rp=%.2e rd=%.2e mu=%.2e delta=%.1e rho=%.1e nfact=%d\n", it, out.rp, out.rd, mu / c, delta, rho_p, out.nfact);
        if (!(mu == mu) || !(out.rd == out.rd)) break;  // NaN guard
        pf.lap(PS_RESID);
        // ---- barrier update: shrink mu_t while the current barrier problem is solved to kappa*mu_t ---
        for (int g = 0; g < 60; ++g) {
            double comp = fmax(fabs(sz[0] - mu_t), fabs(-sz[1] - mu_t));
            double e_mu = fmax(sz[2], comp);
            if (e_mu <= o.ipm_kappa_eps * mu_t && mu_t > o.ipm_mu_min) mu_t = fmax(o.ipm_mu_min, fmin(0.2 * mu_t, mu_t * sqrt(mu_t)));
            else break;
        }
        // ---- weights, assembly, factorisation (with inertia correction) ---------------------------
        for_n(T, M, [&](int i) {
            bool eq = I.mv[M_RL][i] == I.mv[M_RU][i];
            double wi = 0.0;
            if (eq) wi = 1.0 / delta;
            else {
                if (!isinf(I.mv[M_RU][i])) wi += I.mv[M_YC][i] / (I.mv[M_ZC][i] + delta * I.mv[M_YC][i]);
                if (!isinf(I.mv[M_RL][i])) wi += I.mv[M_BC][i] / (I.mv[M_RC][i] + delta * I.mv[M_BC][i]);
            }
            I.mv[M_RW][i] = wi;
        });
        bool fact_ok = false;
        for (int tries = 0; tries < 30 && !fact_ok; ++tries) {
            for_n(T, N, [&](int j) {
                bool eq = I.nv[N_XL][j] == I.nv[N_XU][j];
                double wj = 0.0;
                if (eq) wj = 1.0 / delta;
                else {
                    if (!isinf(I.nv[N_XU][j])) wj += I.nv[N_YB][j] / (I.nv[N_ZB][j] + delta * I.nv[N_YB][j]);
                    if (!isinf(I.nv[N_XL][j])) wj += I.nv[N_KP][j] / (I.nv[N_RB][j] + delta * I.nv[N_KP][j]);
                }
                I.nv[N_DSH][j] = wj + rho_p + (I.useH ? 0.0 : I.nv[N_HD][j]);
            });
            T.sync();
            pf.lap(PS_WEIGHTS);
            chol_assemble(T, C, W, I.useH ? I.Hsv : (const double*)nullptr, I.nv[N_DSH], I.mv[M_RW], I.Jsv);
            pf.lap(PS_ASSEMBLE);
            fact_ok = chol_factor(T, C, W, pf);
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
            // t_row, then I.nv[N_P] = -r_x - T t_row - t_box
            for_n(T, M, [&](int i) {
                bool eq = I.mv[M_RL][i] == I.mv[M_RU][i];
                double ax = I.mv[M_I3][i], ti = 0.0;
                if (eq) ti = I.mv[M_AX][i] / delta;
                else {
                    if (!isinf(I.mv[M_RU][i])) {
                        double rc = sigma_mu - I.mv[M_ZC][i] * I.mv[M_YC][i] - 0.0;
                        ti += (rc + I.mv[M_YC][i] * (I.mv[M_AX][i])) / (I.mv[M_ZC][i] + delta * I.mv[M_YC][i]);
                    }
                    if (!isinf(I.mv[M_RL][i])) {
                        double rc = sigma_mu - I.mv[M_RC][i] * I.mv[M_BC][i] - 0.0;
                        ti -= (rc + I.mv[M_BC][i] * (I.mv[M_I4][i])) / (I.mv[M_RC][i] + delta * I.mv[M_BC][i]);
                    }
                }
                I.mv[M_T][i] = ti;
            });
            T.sync();
            csr_rows(T, N, I.lgT, I.T.rb, I.T.re, I.T.col, I.Tsv, I.mv[M_T], [&](int j, double tt) {
                bool eq = I.nv[N_XL][j] == I.nv[N_XU][j];
                double tb = 0.0;
                if (eq) tb = I.nv[N_TMP2][j] / delta;
                else {
                    if (!isinf(I.nv[N_XU][j])) {
                        double rc = sigma_mu - I.nv[N_ZB][j] * I.nv[N_YB][j] - 0.0;
                        tb += (rc + I.nv[N_YB][j] * (I.nv[N_TMP2][j])) / (I.nv[N_ZB][j] + delta * I.nv[N_YB][j]);
                    }
                    if (!isinf(I.nv[N_XL][j])) {
                        double rc = sigma_mu - I.nv[N_RB][j] * I.nv[N_KP][j] - 0.0;
                        tb -= (rc + I.nv[N_KP][j] * (I.nv[N_I1][j])) / (I.nv[N_RB][j] + delta * I.nv[N_KP][j]);
                    }
                }
                I.nv[N_P][j] = -I.nv[N_R][j] - tt - tb;
            });
            T.sync();
            pf.lap(PS_RHS);
            chol_solve(T, C, W, I.nv[N_P], I.nv[N_XT], pf);
            // iterative refinement against the matrix-free K (K is ill-conditioned by design)
            for (int rf = 0; rf < o.ipm_refine; ++rf) {
                apply_K(T, I, I.nv[N_XT], I.nv[N_TMP], I.nv[N_DSH], I.mv[M_RW], (const double*)nullptr);
                T.sync();
                // apply_K adds I.nv[N_HD] for !useH on top of dsh; I.nv[N_DSH] already contains it -> subtract once
                double nr[2] = {0.0, 0.0};
                for_n(T, N, [&](int j) {
                    double kv = I.nv[N_TMP][j] - (I.useH ? 0.0 : I.nv[N_HD][j] * I.nv[N_XT][j]);
                    double r = I.nv[N_P][j] - kv;
                    I.nv[N_TMP][j] = r;
                    nr[0] = fmax(nr[0], fabs(r));
                    nr[1] = fmax(nr[1], fabs(I.nv[N_P][j]));
                });
                if (o.verbose) {
                    T.template reduce<2, true>(nr);
                    if (T.tid() == 0) printf("      pass %d refine %d: |I.nv[N_P] - K I.nv[N_XT]| = %.2e  |I.nv[N_P]| = %.2e\n", pass, rf, nr[0], nr[1]);
                }
                T.sync();
                chol_solve(T, C, W, I.nv[N_TMP], I.nv[N_TMP], pf);
                for_n(T, N, [&](int j) { I.nv[N_XT][j] += I.nv[N_TMP][j]; });
                T.sync();
            }
            // J I.nv[N_XT] and the step-to-boundary ratio
            double ratio[1] = {0.0};
            csr_rows(T, M, I.lgJ, I.J.rb, I.J.re, I.J.col, I.Jsv, I.nv[N_XT], [&](int i, double jd) {
                I.mv[M_I2][i] = jd;
                if (I.mv[M_RL][i] == I.mv[M_RU][i]) return;
                double ax = I.mv[M_I3][i];
                if (!isinf(I.mv[M_RU][i])) {
                    double rc = sigma_mu - I.mv[M_ZC][i] * I.mv[M_YC][i] - 0.0;
                    SideDir d = side_dir(rc, I.mv[M_YC][i], I.mv[M_ZC][i], I.mv[M_AX][i], jd, delta);
                    ratio[0] = fmax(ratio[0], fmax(step_ratio(I.mv[M_ZC][i], d.ds), step_ratio(I.mv[M_YC][i], d.dz)));
                }
                if (!isinf(I.mv[M_RL][i])) {
                    double rc = sigma_mu - I.mv[M_RC][i] * I.mv[M_BC][i] - 0.0;
                    SideDir d = side_dir(rc, I.mv[M_BC][i], I.mv[M_RC][i], I.mv[M_I4][i], -jd, delta);
                    ratio[0] = fmax(ratio[0], fmax(step_ratio(I.mv[M_RC][i], d.ds), step_ratio(I.mv[M_BC][i], d.dz)));
                }
            });
            for_n(T, N, [&](int j) {
                if (I.nv[N_XL][j] == I.nv[N_XU][j]) return;
                if (!isinf(I.nv[N_XU][j])) {
                    double rc = sigma_mu - I.nv[N_ZB][j] * I.nv[N_YB][j] - 0.0;
                    SideDir d = side_dir(rc, I.nv[N_YB][j], I.nv[N_ZB][j], I.nv[N_TMP2][j], I.nv[N_XT][j], delta);
                    ratio[0] = fmax(ratio[0], fmax(step_ratio(I.nv[N_ZB][j], d.ds), step_ratio(I.nv[N_YB][j], d.dz)));
                }
                if (!isinf(I.nv[N_XL][j])) {
                    double rc = sigma_mu - I.nv[N_RB][j] * I.nv[N_KP][j] - 0.0;
                    SideDir d = side_dir(rc, I.nv[N_KP][j], I.nv[N_RB][j], I.nv[N_I1][j], -I.nv[N_XT][j], delta);
                    ratio[0] = fmax(ratio[0], fmax(step_ratio(I.nv[N_RB][j], d.ds), step_ratio(I.nv[N_KP][j], d.dz)));
                }
            });
            T.template reduce<1, true>(ratio);
            if (pass == 0) {
                double a_aff = ratio[0] > 1.0 ? 1.0 / ratio[0] : 1.0;
                // mu_aff and the second-order cross terms ds_aff * dz_aff
                double ms[1] = {0.0};
                for_n(T, M, [&](int i) {
                    if (I.mv[M_RL][i] == I.mv[M_RU][i]) return;
                    double ax = I.mv[M_I3][i], jd = I.mv[M_I2][i];
                    if (!isinf(I.mv[M_RU][i])) {
                        SideDir d = side_dir(-I.mv[M_ZC][i] * I.mv[M_YC][i], I.mv[M_YC][i], I.mv[M_ZC][i], I.mv[M_AX][i], jd, delta);
                        I.mv[M_YP][i] = d.ds * d.dz;
                        ms[0] += (I.mv[M_ZC][i] + a_aff * d.ds) * (I.mv[M_YC][i] + a_aff * d.dz);
                    }
                    if (!isinf(I.mv[M_RL][i])) {
                        SideDir d = side_dir(-I.mv[M_RC][i] * I.mv[M_BC][i], I.mv[M_BC][i], I.mv[M_RC][i], I.mv[M_I4][i], -jd, delta);
                        I.mv[M_TMP][i] = d.ds * d.dz;
                        ms[0] += (I.mv[M_RC][i] + a_aff * d.ds) * (I.mv[M_BC][i] + a_aff * d.dz);
                    }
                });
                for_n(T, N, [&](int j) {
                    if (I.nv[N_XL][j] == I.nv[N_XU][j]) return;
                    if (!isinf(I.nv[N_XU][j])) {
                        SideDir d = side_dir(-I.nv[N_ZB][j] * I.nv[N_YB][j], I.nv[N_YB][j], I.nv[N_ZB][j], I.nv[N_TMP2][j], I.nv[N_XT][j], delta);
                        I.nv[N_MINV][j] = d.ds * d.dz;
                        ms[0] += (I.nv[N_ZB][j] + a_aff * d.ds) * (I.nv[N_YB][j] + a_aff * d.dz);
                    }
                    if (!isinf(I.nv[N_XL][j])) {
                        SideDir d = side_dir(-I.nv[N_RB][j] * I.nv[N_KP][j], I.nv[N_KP][j], I.nv[N_RB][j], I.nv[N_I1][j], -I.nv[N_XT][j], delta);
                        I.nv[N_XFIX][j] = d.ds * d.dz;
                        ms[0] += (I.nv[N_RB][j] + a_aff * d.ds) * (I.nv[N_KP][j] + a_aff * d.dz);
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
        pf.lap(PS_RATIO);
        // ---- update with the corrector direction -------------------------------------------------------
        for_n(T, M, [&](int i) {
            double ax = I.mv[M_I3][i], jd = I.mv[M_I2][i];
            I.mv[M_I3][i] = ax + alpha * jd;
            if (I.mv[M_RL][i] == I.mv[M_RU][i]) { I.mv[M_I1][i] += alpha * (jd + I.mv[M_AX][i]) / delta; I.mv[M_AX][i] += alpha * jd; return; }
            if (!isinf(I.mv[M_RU][i])) {
                SideDir d = side_dir(sigma_mu - I.mv[M_ZC][i] * I.mv[M_YC][i] , I.mv[M_YC][i], I.mv[M_ZC][i], I.mv[M_AX][i], jd, delta);
                I.mv[M_ZC][i] += alpha * d.ds; I.mv[M_YC][i] += alpha * d.dz; I.mv[M_AX][i] += alpha * (jd + d.ds);
            }
            if (!isinf(I.mv[M_RL][i])) {
                SideDir d = side_dir(sigma_mu - I.mv[M_RC][i] * I.mv[M_BC][i] , I.mv[M_BC][i], I.mv[M_RC][i], I.mv[M_I4][i], -jd, delta);
                I.mv[M_RC][i] += alpha * d.ds; I.mv[M_BC][i] += alpha * d.dz; I.mv[M_I4][i] += alpha * (-jd + d.ds);
            }
        });
        for_n(T, N, [&](int j) {
            double xj = I.nv[N_X][j], dj = I.nv[N_XT][j];
            if (I.nv[N_XL][j] == I.nv[N_XU][j]) { I.nv[N_MASK][j] += alpha * (dj + I.nv[N_TMP2][j]) / delta; I.nv[N_TMP2][j] += alpha * dj; I.nv[N_X][j] = xj + alpha * dj; return; }
            if (!isinf(I.nv[N_XU][j])) {
                SideDir d = side_dir(sigma_mu - I.nv[N_ZB][j] * I.nv[N_YB][j] , I.nv[N_YB][j], I.nv[N_ZB][j], I.nv[N_TMP2][j], dj, delta);
                I.nv[N_ZB][j] += alpha * d.ds; I.nv[N_YB][j] += alpha * d.dz; I.nv[N_TMP2][j] += alpha * (dj + d.ds);
            }
            if (!isinf(I.nv[N_XL][j])) {
                SideDir d = side_dir(sigma_mu - I.nv[N_RB][j] * I.nv[N_KP][j] , I.nv[N_KP][j], I.nv[N_RB][j], I.nv[N_I1][j], -dj, delta);
                I.nv[N_RB][j] += alpha * d.ds; I.nv[N_KP][j] += alpha * d.dz; I.nv[N_I1][j] += alpha * (-dj + d.ds);
            }
            I.nv[N_X][j] = xj + alpha * dj;
        });
        pf.lap(PS_UPDATE);
        delta = fmax(o.ipm_delta_min, delta * 0.3);
        if (rho_p > o.ipm_rho0) rho_p = fmax(o.ipm_rho0, rho_p / 3.0);
        out.iters = it + 1;
    }
    T.sync();
    if (!out.solved && acc_cnt > 0) out.solved = true;  // iteration cap reached on an acceptable point
    if (out.solved || out.almost) {
        // multipliers in the OSQP sign the output stage expects: yc (rows) and yb (box)
        double *yc = I.mv[M_YC], *yb = I.nv[N_YB];
        for_n(T, M, [&](int i) { yc[i] = (I.mv[M_RL][i] == I.mv[M_RU][i]) ? I.mv[M_I1][i] : (I.mv[M_YC][i] - I.mv[M_BC][i]); });
        for_n(T, N, [&](int j) { yb[j] = (I.nv[N_XL][j] == I.nv[N_XU][j]) ? I.nv[N_MASK][j] : (I.nv[N_YB][j] - I.nv[N_KP][j]); });
        T.sync();
    }
    return out;
}
