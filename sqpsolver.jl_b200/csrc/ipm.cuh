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
//   monotone (Fiacco-McCormick) barrier rule, fraction-to-the-boundary step, delta -> delta_min geometrically
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

// assembly + factorisation and the Newton solve, by team.  The resident CTA team streams the ring program when it has one:
// the right-hand side b is then needed BEFORE the factorisation, because its forward sweep rides in the factor chunks.
template <class Team>
__device__ __forceinline__ bool team_assemble_factor(Team& T, Ring*, const CholDev& C, const CholWork& W, const double* Pv, const double* dg,
                                                     double shift, const double* w, const double* Jv, const double* b, Prof& pf) {
    chol_assemble(T, C, W, Pv, dg, shift, w, Jv, pf);
    return chol_factor(T, C, W, pf, b);
}
__device__ __forceinline__ bool team_assemble_factor(CtaTeam& T, Ring* R, const CholDev& C, const CholWork& W, const double* Pv,
                                                     const double* dg, double shift, const double* w, const double* Jv,
                                                     const double* b, Prof& pf) {
    if (R) return chol_assemble_factor_fwd_ring(T, *R, C, W, Pv, dg, shift, w, Jv, b, pf);
    chol_assemble(T, C, W, Pv, dg, shift, w, Jv, pf);
    return chol_factor(T, C, W, pf, b);
}
template <class Team>
__device__ __forceinline__ void team_solve(Team& T, Ring*, const CholDev& C, const CholWork& W, const double* b, double* x, Prof& pf) {
    chol_solve(T, C, W, b, x, pf, C.fused_fwd != 0);
}
__device__ __forceinline__ void team_solve(CtaTeam& T, Ring* R, const CholDev& C, const CholWork& W, const double* b, double* x, Prof& pf) {
    if (R) chol_backsolve_ring(T, *R, C, W, x, pf);
    else chol_solve(T, C, W, b, x, pf, C.fused_fwd != 0);
}

template <bool RING, class Team>
__device__ IpmOut ipm_run(Team& T, const Inst& I, const CholDev& C, const CholWork& W, Ring* Rin, const sqpqp_options& o,
                          double c, int phase, const double* xk_scaled_start, int handoff_k, IpmState* st, bool resume) {
    Ring* const R = RING ? Rin : (Ring*)nullptr;  // a compile-time null outside the resident launch: the ring code folds away
    const int N = I.N, M = I.M;
    // Work vectors are slots of the ADMM workspace (the two methods never run concurrently), addressed through
    // I.nv / I.mv at the point of use (kernel-parameter bank, no per-thread pointer table):
    //   rows: s_u M_ZC, z_u M_YC, s_l M_RC, z_l M_BC, y M_I1, tracked residuals r_u M_AX / r_l M_I4, A x M_I3,
    //         J dx M_I2, w M_RW, t M_T, rhs coefficients a M_YP / b M_TMP
    //   cols: x N_X, s_u N_ZB, z_u N_YB, s_l N_RB, z_l N_KP, y N_MASK, r_u N_TMP2 / r_l N_I1, r_x N_R, rhs N_P,
    //         dx N_XT, lam_box N_TMP, diagonal of K N_DSH, a N_MINV / b N_XFIX
    // Side residuals are TRACKED (r += alpha (g dx + ds)), never recomputed from A x - b: at the end of the
    // solve delta ~ 1e-8 and a freshly evaluated residual carries ~1e-16 |Ax| of rounding noise, which the
    // dual update dy = (J dx + r)/delta would amplify by 1/delta.
    IpmOut out{false, false, false, false, false, 0, 0, INFINITY, INFINITY, 0.0};
    Prof pf;
    pf.start();
    double cnt[1] = {0.0};
    if (resume) {
        // continue an instance another launch handed over: its iterate lives in the global work vectors; bring the parts this
        // launch keeps in shared memory there (the scaled problem data was rebuilt, identically, by stage A)
        const int stN[] = {N_X, N_ZB, N_YB, N_RB, N_KP, N_MASK, N_TMP2, N_I1, N_XT};
        const int stM[] = {M_ZC, M_YC, M_RC, M_BC, M_I1, M_AX, M_I4, M_I3, M_I2};
        for (int q = 0; q < 9; ++q) {
            double* dst = I.nv[stN[q]];
            const double* src = I.nv.global(stN[q]);
            if (dst != src) for_n(T, N, [&](int j) { dst[j] = src[j]; });
            double* dm = I.mv[stM[q]];
            const double* sm = I.mv.global(stM[q]);
            if (dm != sm) for_n(T, M, [&](int r) { dm[r] = sm[r]; });
        }
        T.sync();
    } else {
    // ---- start point: I.nv[N_X] inside the box, unit duals, slacks >= 1 --------------------------------
    for_n(T, N, [&](int j) {
        double v = xk_scaled_start ? xk_scaled_start[j] : 0.0;
        v = fmin(fmax(v, I.nv[N_XL][j]), I.nv[N_XU][j]);
        I.nv[N_X][j] = v;
        I.nv[N_MASK][j] = 0.0;
    });
    T.sync();
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
    }
    const double nin = resume ? st->nin : fmax(cnt[0], 1.0);

    double delta = o.ipm_delta0, rho_p = o.ipm_rho0, rho_last = 0.0;
    int acc_cnt = 0;
    double rp_ref = INFINITY;
    double mu_t = o.ipm_mu0;  // barrier parameter of the current subproblem (monotone Fiacco-McCormick rule)
    // The iteration is organised in FUSED passes (every pass loads all the operands of an element first and then
    // computes branch-free on registers, so that a thread has one L2 round trip per element, not one per branch):
    //   P1  rows / cols : apply the previous step (alpha, sigma*mu and delta of the previous solve), then the
    //                     multiplier, the residual norms, the weights w and the right-hand-side coefficients
    //                     a, b  (t = sigma*mu * a + b) for the new delta
    //   P2  cols (CSR)  : r_x = P x + q + J' lam + lam_box and its norms
    //   -- reductions, termination tests, barrier update, assembly + factorisation (inertia correction) --
    //   P3  rows        : t = sigma*mu * a + b
    //   P4  cols (CSR)  : rhs = -r_x - J' t - t_box;  Cholesky solve
    //   P5  rows (CSR) / cols : J dx, step-to-boundary ratio
    double alpha = 0.0, sig_prev = 0.0, del_prev = delta;
    bool capped = true;  // cleared by every exit of the loop other than the iteration cap
    int it0 = 0;
    if (resume) {
        delta = st->delta; rho_p = st->rho_p; rho_last = st->rho_last; mu_t = st->mu_t; alpha = st->alpha;
        sig_prev = st->sig_prev; del_prev = st->del_prev; rp_ref = st->rp_ref; acc_cnt = st->acc_cnt;
        it0 = st->it; out.nfact = st->nfact; out.iters = it0;
    } else {
        for_n(T, M, [&](int i) { I.mv[M_I2][i] = 0.0; });
        for_n(T, N, [&](int j) { I.nv[N_XT][j] = 0.0; });
    }
    T.sync();
    for (int it = it0; it < o.ipm_max_iter; ++it) {
        out.iters = it;
        double mx[8] = {0, 0, 0, 0, 0, 0, 0.0, -INFINITY};
        // [0] rp [1] rd*c [2] primal scale [3] dual scale*c [4] |A'lam|*c [5] |lam|*c (unscaled) [6] max s*z [7] max -(s*z)
        double m2[2] = {0.0, 0.0};   // [0] scaled residual of the barrier problem  [1] max |row multiplier|
        double sums[2] = {0.0, 0.0}; // [0] support function u'(lam)+ + l'(lam)- (scaled units)  [1] sum s*z
        // ---- P1 rows ---------------------------------------------------------------------------
        for_n(T, M, [&](int i) {
            const double rl_ = I.mv[M_RL][i], ru_ = I.mv[M_RU][i], es = I.mv[M_ES][i];
            double su = I.mv[M_ZC][i], zu = I.mv[M_YC][i], sl = I.mv[M_RC][i], zl = I.mv[M_BC][i];
            double yy = I.mv[M_I1][i], rU = I.mv[M_AX][i], rL = I.mv[M_I4][i], ax = I.mv[M_I3][i];
            const double jd = I.mv[M_I2][i];
            const bool eq = rl_ == ru_, uf = !eq && !isinf(ru_), lf = !eq && !isinf(rl_);
            ax += alpha * jd;
            if (eq) { yy += alpha * (jd + rU) / del_prev; rU += alpha * jd; }
            if (uf) {
                SideDir d = side_dir(sig_prev - su * zu, zu, su, rU, jd, del_prev);
                su += alpha * d.ds; zu += alpha * d.dz; rU += alpha * (jd + d.ds);
            }
            if (lf) {
                SideDir d = side_dir(sig_prev - sl * zl, zl, sl, rL, -jd, del_prev);
                sl += alpha * d.ds; zl += alpha * d.dz; rL += alpha * (-jd + d.ds);
            }
            const double lam = eq ? yy : (zu - zl);
            double pr = eq ? fabs(rU) : 0.0, wi = eq ? 1.0 / delta : 0.0, ai = 0.0, bi = eq ? rU / delta : 0.0;
            if (uf) {
                const double p_ = su * zu, dd = 1.0 / (su + delta * zu);
                pr = fmax(pr, fabs(rU)); sums[1] += p_; mx[6] = fmax(mx[6], p_); mx[7] = fmax(mx[7], -p_);
                wi += zu * dd; ai += dd; bi += (zu * rU - p_) * dd;
            }
            if (lf) {
                const double p_ = sl * zl, dd = 1.0 / (sl + delta * zl);
                pr = fmax(pr, fabs(rL)); sums[1] += p_; mx[6] = fmax(mx[6], p_); mx[7] = fmax(mx[7], -p_);
                wi += zl * dd; ai -= dd; bi -= (zl * rL - p_) * dd;
            }
            m2[0] = fmax(m2[0], pr);
            m2[1] = fmax(m2[1], fabs(lam));
            mx[0] = fmax(mx[0], pr / es);
            mx[2] = fmax(mx[2], fabs(ax) / es);
            mx[5] = fmax(mx[5], fabs(lam) * es);
            if (lam > 0.0) sums[0] += ru_ * lam; else if (lam < 0.0) sums[0] += rl_ * lam;
            I.mv[M_ZC][i] = su; I.mv[M_YC][i] = zu; I.mv[M_RC][i] = sl; I.mv[M_BC][i] = zl;
            I.mv[M_I1][i] = yy; I.mv[M_AX][i] = rU; I.mv[M_I4][i] = rL; I.mv[M_I3][i] = ax;
            I.mv[M_T][i] = lam; I.mv[M_RW][i] = wi; I.mv[M_YP][i] = ai; I.mv[M_TMP][i] = bi;
        });
        // ---- P1 cols ---------------------------------------------------------------------------
        for_n(T, N, [&](int j) {
            const double xl_ = I.nv[N_XL][j], xu_ = I.nv[N_XU][j], Dj = I.nv[N_D][j];
            double su = I.nv[N_ZB][j], zu = I.nv[N_YB][j], sl = I.nv[N_RB][j], zl = I.nv[N_KP][j];
            double yy = I.nv[N_MASK][j], rU = I.nv[N_TMP2][j], rL = I.nv[N_I1][j], xj = I.nv[N_X][j];
            const double dj = I.nv[N_XT][j];
            const bool eq = xl_ == xu_, uf = !eq && !isinf(xu_), lf = !eq && !isinf(xl_);
            xj += alpha * dj;
            if (eq) { yy += alpha * (dj + rU) / del_prev; rU += alpha * dj; }
            if (uf) {
                SideDir d = side_dir(sig_prev - su * zu, zu, su, rU, dj, del_prev);
                su += alpha * d.ds; zu += alpha * d.dz; rU += alpha * (dj + d.ds);
            }
            if (lf) {
                SideDir d = side_dir(sig_prev - sl * zl, zl, sl, rL, -dj, del_prev);
                sl += alpha * d.ds; zl += alpha * d.dz; rL += alpha * (-dj + d.ds);
            }
            const double lamb = eq ? yy : (zu - zl);
            double pr = eq ? fabs(rU) : 0.0, wj = eq ? 1.0 / delta : 0.0, aj = 0.0, bj = eq ? rU / delta : 0.0;
            if (uf) {
                const double p_ = su * zu, dd = 1.0 / (su + delta * zu);
                pr = fmax(pr, fabs(rU)); sums[1] += p_; mx[6] = fmax(mx[6], p_); mx[7] = fmax(mx[7], -p_);
                wj += zu * dd; aj += dd; bj += (zu * rU - p_) * dd;
            }
            if (lf) {
                const double p_ = sl * zl, dd = 1.0 / (sl + delta * zl);
                pr = fmax(pr, fabs(rL)); sums[1] += p_; mx[6] = fmax(mx[6], p_); mx[7] = fmax(mx[7], -p_);
                wj += zl * dd; aj -= dd; bj -= (zl * rL - p_) * dd;
            }
            m2[0] = fmax(m2[0], pr);
            mx[0] = fmax(mx[0], pr * Dj);
            mx[2] = fmax(mx[2], fabs(xj) * Dj);
            mx[5] = fmax(mx[5], fabs(lamb) / Dj);
            if (lamb > 0.0) sums[0] += xu_ * lamb; else if (lamb < 0.0) sums[0] += xl_ * lamb;
            I.nv[N_ZB][j] = su; I.nv[N_YB][j] = zu; I.nv[N_RB][j] = sl; I.nv[N_KP][j] = zl;
            I.nv[N_MASK][j] = yy; I.nv[N_TMP2][j] = rU; I.nv[N_I1][j] = rL; I.nv[N_X][j] = xj;
            I.nv[N_TMP][j] = lamb; I.nv[N_DSH][j] = wj + (I.useH ? 0.0 : I.nv[N_HD][j]);
            I.nv[N_MINV][j] = aj; I.nv[N_XFIX][j] = bj;
        });
        T.sync();
        // ---- P2: stationarity residual ---------------------------------------------------------------
        csr_rows2(T, N, I.lgT, I.H.rb, I.H.re, I.H.col, I.Hsv, I.nv[N_X], I.useH, I.T.rb, I.T.re, I.T.col, I.Tsv, I.mv[M_T],
                  [&](int j, double px, double aty) {
                      const double lamb = I.nv[N_TMP][j], qj = I.nv[N_Q][j], id = 1.0 / I.nv[N_D][j];
                      if (!I.useH) px = I.nv[N_HD][j] * I.nv[N_X][j];
                      const double r = px + qj + aty + lamb;
                      I.nv[N_R][j] = r;
                      mx[4] = fmax(mx[4], fabs(aty + lamb) * id);
                      m2[0] = fmax(m2[0], fabs(r));
                      mx[1] = fmax(mx[1], fabs(r) * id);
                      mx[3] = fmax(mx[3], fmax(fabs(px), fmax(fabs(aty), fabs(qj))) * id);
                  });
        T.template reduce<8, true>(mx);
        T.template reduce<2, true>(m2);
        T.template reduce<2, false>(sums);
        pf.lap(PS_RESID);
        const double ymx = m2[1], sup = sums[0];
        if (ymx > 1e12) { out.blowup = true; out.almost = false; capped = false; break; }  // multipliers exploding: ADMM certifies infeasibility
        const double mu = sums[1] / nin;
        out.rp = mx[0];
        out.rd = mx[1] / c;
        const double scale_p = fmax(1.0, mx[2]), scale_d = fmax(1.0, mx[3] / c);
        // primal infeasibility certificate on the multiplier direction (same test as the ADMM path applies
        // to its dual increments): A' lam ~ 0 while the support function of the bounds is negative
        if (mx[5] / c > 1e4 && mx[4] <= o.eps_inf * mx[5] && sup <= -o.eps_inf * mx[5]) {
            out.infeasible = true;
            capped = false;
            break;
        }
        // ... or, for marginally infeasible rows (violation << 1, multipliers growing only linearly):
        // the regularised equalities make the iterates converge to a minimiser of the violation, so a
        // primal residual that has stalled at a positive value while the dual residual is converged is
        // the interior-point analogue of Ipopt's failed restoration phase -> LOCALLY_INFEASIBLE
        if (it >= 20 && it % 10 == 0) {
            bool stalled = out.rp > 1e4 * o.ipm_eps * scale_p && fabs(out.rp - rp_ref) <= 1e-3 * out.rp &&
                           out.rd <= 1e-5 * scale_d && mx[5] / c > 1e3;
            if (stalled) { out.infeasible = true; capped = false; break; }
        }
        if (it % 10 == 0) rp_ref = out.rp;
        // Termination (Ipopt-style scaling): primal residual relative to |x|,|Ax|; stationarity relative to
        // the gradient terms (its attainable floor is ~1e-9 of them, the conditioning of K); complementarity
        // (largest s*z, unscaled) absolute unless the multipliers themselves are large.
        // The start-point projection (LP phase) is solved once per SQP run and its solution IS the first iterate:
        // weakly active bounds are only located to sqrt(s*z), so its complementarity target is 1000x tighter.
        const double comp_u = mx[6] / c, sc = fmax(1.0, ymx / c / 100.0);
        const double eps_c = (phase == SQPQP_PHASE_LP) ? 1e-3 * o.ipm_eps : o.ipm_eps;
        if (out.rp <= o.ipm_eps * scale_p && out.rd <= o.ipm_eps * scale_d && comp_u <= eps_c * sc) {
            out.solved = true;
            capped = false;
            break;
        }
        const double acc_eps = 100.0 * o.ipm_eps;
        bool acceptable = out.rp <= acc_eps * scale_p && out.rd <= acc_eps * scale_d && comp_u <= 100.0 * eps_c * sc;
        acc_cnt = acceptable ? acc_cnt + 1 : 0;
        out.almost = out.rp <= 1e-6 * scale_p && out.rd <= 1e-6 * scale_d && comp_u <= 1e-6 * sc;
        if (acc_cnt >= 8) {  // stuck on the floor of an acceptable point (Ipopt: "solved to acceptable level")
            out.almost = true;
            capped = false;
            break;
        }
        if (o.verbose && T.tid() == 0)
            printf("  ipm %3d rp=%.2e rd=%.2e mu=%.2e delta=%.1e rho=%.1e nfact=%d alpha=%.3e\n", it, out.rp, out.rd, mu / c, delta, rho_p,
                   out.nfact, alpha);
        if (!(mu == mu) || !(out.rd == out.rd)) { out.almost = false; capped = false; break; }  // NaN guard
        // ---- barrier update: shrink mu_t while the current barrier problem is solved to kappa*mu_t ---
        for (int g = 0; g < 60; ++g) {
            double comp = fmax(fabs(mx[6] - mu_t), fabs(-mx[7] - mu_t));
            double e_mu = fmax(m2[0], comp);
            if (e_mu <= o.ipm_kappa_eps * mu_t && mu_t > o.ipm_mu_min) mu_t = fmax(o.ipm_mu_min, fmin(0.2 * mu_t, mu_t * sqrt(mu_t)));
            else break;
        }
        pf.lap(PS_WEIGHTS);
        const double sigma_mu = mu_t;
        const double tau_k = fmax(o.ipm_tau, 1.0 - mu_t);
        auto build_rhs = [&]() {  // P3 / P4: rhs = -r_x - J' t - t_box with t = sigma*mu * a + b
            for_n(T, M, [&](int i) { I.mv[M_T][i] = fma(sigma_mu, I.mv[M_YP][i], I.mv[M_TMP][i]); });
            T.sync();
            csr_rows(T, N, I.lgT, I.T.rb, I.T.re, I.T.col, I.Tsv, I.mv[M_T], [&](int j, double tt) {
                const double tb = fma(sigma_mu, I.nv[N_MINV][j], I.nv[N_XFIX][j]);
                I.nv[N_P][j] = -I.nv[N_R][j] - tt - tb;
            });
            T.sync();
            pf.lap(PS_RHS);
        };
        const bool rhs_first = R || C.fused_fwd;  // the forward sweep of the Newton solve rides in the factorisation's phases
        if (rhs_first) build_rhs();
        // ---- assembly, factorisation (with inertia correction: the shift rho_p is a scalar on the diagonal) ----
        bool fact_ok = false;
        for (int tries = 0; tries < 30 && !fact_ok; ++tries) {
            fact_ok = team_assemble_factor(T, R, C, W, I.useH ? I.Hsv : (const double*)nullptr, I.nv[N_DSH], rho_p, I.mv[M_RW], I.Jsv, I.nv[N_P], pf);
            ++out.nfact;
            // Growth 4 on a failed factorisation (the shift decays by 3 per iteration, so this returns to just above the
            // last value that worked).  The textbook 8-10 overshoots on the indefinite ACOPF subproblems: the extra
            // regularisation shortens the Newton steps, the shift decays, fails and overshoots again -- the six slowest
            // recorded subproblems need 228 iterations / 330 factorisations in total with 4 against 504 / 712 with 10
            // (profiles/r01_tuning.md section 6).
            if (!fact_ok) rho_p = fmax(fmax(o.ipm_ic_growth * rho_p, rho_last > 0.0 ? rho_last / o.ipm_ic_decay : 1e-4), 1e-6);
            if (rho_p > 1e8) break;
        }
        if (!fact_ok) { out.almost = false; capped = false; break; }
        if (rho_p > 10.0 * o.ipm_rho0) rho_last = rho_p;
        out.rho_p = rho_p;

        // ---- P3 / P4: right-hand side, Newton solve ----------------------------------------------------
        if (!rhs_first) build_rhs();
        team_solve(T, R, C, W, I.nv[N_P], I.nv[N_XT], pf);
        // ---- P5: J dx and the step-to-boundary ratio ------------------------------------------------------
        double ratio[1] = {0.0};
        csr_rows(T, M, I.lgJ, I.J.rb, I.J.re, I.J.col, I.Jsv, I.nv[N_XT], [&](int i, double jd) {
            const double rl_ = I.mv[M_RL][i], ru_ = I.mv[M_RU][i];
            const double su = I.mv[M_ZC][i], zu = I.mv[M_YC][i], sl = I.mv[M_RC][i], zl = I.mv[M_BC][i];
            const double rU = I.mv[M_AX][i], rL = I.mv[M_I4][i];
            I.mv[M_I2][i] = jd;
            const bool eq = rl_ == ru_;
            if (!eq && !isinf(ru_)) {
                SideDir d = side_dir(sigma_mu - su * zu, zu, su, rU, jd, delta);
                ratio[0] = fmax(ratio[0], fmax(step_ratio(su, d.ds), step_ratio(zu, d.dz)));
            }
            if (!eq && !isinf(rl_)) {
                SideDir d = side_dir(sigma_mu - sl * zl, zl, sl, rL, -jd, delta);
                ratio[0] = fmax(ratio[0], fmax(step_ratio(sl, d.ds), step_ratio(zl, d.dz)));
            }
        });
        for_n(T, N, [&](int j) {
            const double xl_ = I.nv[N_XL][j], xu_ = I.nv[N_XU][j];
            const double su = I.nv[N_ZB][j], zu = I.nv[N_YB][j], sl = I.nv[N_RB][j], zl = I.nv[N_KP][j];
            const double rU = I.nv[N_TMP2][j], rL = I.nv[N_I1][j], dj = I.nv[N_XT][j];
            const bool eq = xl_ == xu_;
            if (!eq && !isinf(xu_)) {
                SideDir d = side_dir(sigma_mu - su * zu, zu, su, rU, dj, delta);
                ratio[0] = fmax(ratio[0], fmax(step_ratio(su, d.ds), step_ratio(zu, d.dz)));
            }
            if (!eq && !isinf(xl_)) {
                SideDir d = side_dir(sigma_mu - sl * zl, zl, sl, rL, -dj, delta);
                ratio[0] = fmax(ratio[0], fmax(step_ratio(sl, d.ds), step_ratio(zl, d.dz)));
            }
        });
        T.template reduce<1, true>(ratio);
        alpha = 1.0;
        if (ratio[0] > 0.0) alpha = fmin(1.0, tau_k / ratio[0]);
        pf.lap(PS_RATIO);
        // the step itself is applied by pass P1 of the next iteration, with the sigma*mu and delta of THIS solve
        sig_prev = sigma_mu;
        del_prev = delta;
        delta = fmax(o.ipm_delta_min, delta * 0.3);
        if (rho_p > o.ipm_rho0) rho_p = fmax(o.ipm_rho0, rho_p / o.ipm_ic_decay);
        out.iters = it + 1;
        if (handoff_k > 0 && it + 1 >= handoff_k && it + 1 < o.ipm_max_iter) {
            // iteration quota of the throughput launch used up: the loop state goes to global memory (the vectors are there
            // already: this launch keeps no work vector in shared memory) and the resident launch takes over at iteration it + 1
            if (T.tid() == 0) {
                st->delta = delta; st->rho_p = rho_p; st->rho_last = rho_last; st->mu_t = mu_t; st->alpha = alpha;
                st->sig_prev = sig_prev; st->del_prev = del_prev; st->rp_ref = rp_ref; st->nin = nin;
                st->it = it + 1; st->nfact = out.nfact; st->acc_cnt = acc_cnt;
            }
            out.handoff = true;
            capped = false;
            break;
        }
    }
    T.sync();
    if (R) ring_drain(*R);  // chunks requested ahead for an iteration that is not going to happen
    if (out.handoff) return out;
    // Iteration cap reached while the iterate was at the acceptable level (100 x ipm_eps on every residual): returned as
    // ALMOST_LOCALLY_SOLVED, Ipopt's "solved to acceptable level".  Exits through a blow-up, a NaN or a failed
    // factorisation never promote an iterate (their acc_cnt belongs to an earlier iteration).
    if (!out.solved && capped && acc_cnt > 0) out.almost = true;
    if (out.solved || out.almost) {
        // multipliers in the OSQP sign the output stage expects: yc (rows) and yb (box)
        double *yc = I.mv[M_YC], *yb = I.nv[N_YB];
        for_n(T, M, [&](int i) { yc[i] = (I.mv[M_RL][i] == I.mv[M_RU][i]) ? I.mv[M_I1][i] : (I.mv[M_YC][i] - I.mv[M_BC][i]); });
        for_n(T, N, [&](int j) { yb[j] = (I.nv[N_XL][j] == I.nv[N_XU][j]) ? I.nv[N_MASK][j] : (I.nv[N_YB][j] - I.nv[N_KP][j]); });
        T.sync();
    }
    return out;
}
