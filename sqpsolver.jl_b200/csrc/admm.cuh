// admm.cuh -- the QP-subproblem solve, entirely on device (one kernel launch per batch).
//
// Replaces the external solve behind JuMP.optimize! in sub_optimize! /
// sub_optimize_FR! / sub_optimize_lp (subproblem_JuMP.jl:178, 388, 209) together with
// set_trust_region! (:432-448), modify_constraints! (:465-512) and
// collect_solution! (:514-563).
//
//   min 1/2 x'Px + q'x   s.t.  rl <= Jx <= ru,   xl <= x <= xu          (OSQP form, A = [J; I])
//
// Steps (all inside solve_instance, no host round trips):
//   0. assemble q, rl, ru, xl, xu for the phase (QP / FR / SOC / LP projection)
//   1. Ruiz equilibration of [P J'; J 0] + cost scaling; scaled copies of the values
//   2. if P may be indefinite: lambda_min(P_scaled) by shifted power iteration ->
//      floor on the box step size rho_b (nonconvexity guard)
//   3. ADMM with relaxation alpha; x-update by Jacobi-PCG on
//         K = P + sigma I + diag(rho_b) + J' diag(rho_c) J         (matrix-free: 3 SpMV)
//      every `check_every` iterations: unscaled residuals, primal-infeasibility
//      certificate, rho adaptation, active-set prediction
//   4. polish: predicted-active rows as equalities (method of multipliers with the same
//      PCG), active bounds eliminated; accepted only as a verified KKT point
//   5. unscale, map to MOI signs and the reference's storage convention.
#pragma once
#include "team.cuh"

struct DevOpts {
    sqpqp_options o;
    int handoff_k;   // > 0: an interior-point solve that has not finished after this many iterations saves its state and is
                     //      flagged 3 for the resident launch that follows (stragglers finish there at lower latency)
    int resume;      // 1: this launch continues the instances flagged 3 and skips all others
    const int* order;  // nullable: CTA slot k of the launch runs instance order[k] (longest-first order of a resumed batch)
};

// A work-vector slot resolved to its address at the point of use, from kernel parameters (constant
// bank) only: base[k] + instance offset, or the CTA's shared-memory copy when the slot is resident.
// (A per-thread table of the ~40 resolved pointers does not fit the 64-register budget of the
// 4-CTAs-per-SM build: it lived in local memory and its reloads were 20 % of all load instructions
// and the larger part of the kernel's DRAM traffic -- profiles/r01_ncu_summary.md.)
struct VecTab {
    double* const* base;
    const int* off;   // shared-memory offsets (doubles) or nullptr
    double* dsm;
    size_t inst_off;
    __device__ __forceinline__ double* operator[](int k) const {
        if (off) {
            int o = off[k];
            if (o >= 0) return dsm + o;
        }
        return base[k] + inst_off;
    }
    __device__ __forceinline__ double* global(int k) const { return base[k] + inst_off; }  // the slot's home in global memory
};

// resolved per-instance views
struct Inst {
    int N, M;            // active columns / rows this phase
    int n, m, S;
    bool useH;
    double pconst;       // P = useH*H + pconst*I
    Csr J, T, H;
    int lgJ, lgT, lgH;
    const double *Jv, *Tv, *Hv;
    double *Jsv, *Tsv, *Hsv;
    VecTab nv, mv;
    signed char *codeC, *codeB, *prevC, *prevB, *triedC, *triedB;
};

template <class Team, class F>
__device__ __forceinline__ void for_n(Team& T, int len, F f) {
    for (int i = T.tid(); i < len; i += T.size()) f(i);
}

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

// ---- K v = Ps v + dsh.*v + Ts (rw .* (Js v)), masked; returns partial of v'Kv ---------
template <class Team>
__device__ __forceinline__ double apply_K(Team& T, const Inst& I, const double* __restrict__ v, double* __restrict__ out,
                                          const double* __restrict__ dsh, const double* __restrict__ rw,
                                          const double* __restrict__ mask) {
    double* t = I.mv[M_T];
    csr_rows(T, I.M, I.lgJ, I.J.rb, I.J.re, I.J.col, I.Jsv, v, [&](int r, double d) { t[r] = rw[r] * d; });
    T.sync();
    double part = 0.0;
    const double* hd = I.nv[N_HD];
    csr_rows2(T, I.N, I.lgT, I.H.rb, I.H.re, I.H.col, I.Hsv, v, I.useH, I.T.rb, I.T.re, I.T.col, I.Tsv, t,
              [&](int r, double a1, double a2) {
                  double k = a1 + a2 + (dsh[r] + (I.useH ? 0.0 : hd[r])) * v[r];
                  if (mask) k *= mask[r];
                  out[r] = k;
                  part = fma(v[r], k, part);
              });
    return part;
}

struct CgOut {
    int iters;
    bool neg;
    double curv;  // Rayleigh quotient at breakdown
};

// Jacobi-preconditioned CG.  On entry: b in N_R, start in xs.  mask may be null.
template <class Team>
__device__ CgOut pcg(Team& T, const Inst& I, double* __restrict__ xs, const double* __restrict__ dsh,
                     const double* __restrict__ rw, const double* __restrict__ mask, double tol_abs, double rel0,
                     int maxit) {
    double *r = I.nv[N_R], *p = I.nv[N_P], *Kp = I.nv[N_KP];
    const double* Mi = I.nv[N_MINV];
    CgOut out{0, false, 0.0};
    T.sync();
    apply_K(T, I, xs, Kp, dsh, rw, mask);
    T.sync();
    double red[2] = {0.0, 0.0};
    for_n(T, I.N, [&](int i) {
        double ri = r[i] - Kp[i];
        if (mask) ri *= mask[i];
        r[i] = ri;
        double z = Mi[i] * ri;
        p[i] = z;
        red[0] = fma(ri, z, red[0]);
        red[1] = fma(ri, ri, red[1]);
    });
    T.template reduce<2, false>(red);
    double rz = red[0], rr = red[1];
    // stop at max(tol_abs, rel0 * |r0|): the warm-started residual r0 is the change of the
    // right-hand side since the previous ADMM iteration, so this is a relative-error criterion
    tol_abs = fmax(tol_abs, rel0 * sqrt(rr));
    while (rr > tol_abs * tol_abs && out.iters < maxit) {
        T.sync();  // p complete before the gathers
        double q[2];
        q[0] = apply_K(T, I, p, Kp, dsh, rw, mask);
        q[1] = 0.0;
        for_n(T, I.N, [&](int i) { q[1] = fma(p[i], p[i], q[1]); });
        T.template reduce<2, false>(q);
        if (!(q[0] > 0.0)) {
            out.neg = true;
            out.curv = q[1] > 0.0 ? q[0] / q[1] : 0.0;
            break;
        }
        double a = rz / q[0];
        red[0] = red[1] = 0.0;
        for_n(T, I.N, [&](int i) {
            xs[i] = fma(a, p[i], xs[i]);
            double ri = fma(-a, Kp[i], r[i]);
            r[i] = ri;
            double z = Mi[i] * ri;
            red[0] = fma(ri, z, red[0]);
            red[1] = fma(ri, ri, red[1]);
        });
        T.template reduce<2, false>(red);
        double beta = red[0] / rz;
        rz = red[0];
        rr = red[1];
        for_n(T, I.N, [&](int i) { p[i] = fma(beta, p[i], Mi[i] * r[i]); });
        ++out.iters;
    }
    T.sync();
    return out;
}

// Minv = 1 / (diag(Ps) + dsh + sum_i rw_i Js_ij^2)
template <class Team>
__device__ void build_minv(Team& T, const Inst& I, const double* __restrict__ dsh, const double* __restrict__ rw) {
    double* Mi = I.nv[N_MINV];
    const double* hd = I.nv[N_HD];
    for (int j = T.tid(); j < I.N; j += T.size()) {
        double d = hd[j] + dsh[j];
        for (int k = I.T.rb[j]; k < I.T.re[j]; ++k) {
            double a = I.Tsv[k];
            d = fma(rw[I.T.col[k]] * a, a, d);
        }
        Mi[j] = 1.0 / fmax(d, 1e-300);
    }
    T.sync();
}

template <class Team>
__device__ void set_rho(Team& T, const Inst& I, const sqpqp_options& o, double rho, double rb_floor) {
    double *rc = I.mv[M_RC], *rb = I.nv[N_RB];
    const double *rl = I.mv[M_RL], *ru = I.mv[M_RU], *xl = I.nv[N_XL], *xu = I.nv[N_XU];
    for_n(T, I.M, [&](int i) {
        double v = rho;
        if (rl[i] == ru[i]) v = rho * o.rho_eq_mult;
        else if (isinf(rl[i]) && isinf(ru[i])) v = o.rho_min;
        rc[i] = v;
    });
    for_n(T, I.N, [&](int j) {
        double v = rho;
        if (xl[j] == xu[j]) v = rho * o.rho_eq_mult;
        else if (isinf(xl[j]) && isinf(xu[j])) v = o.rho_min;
        rb[j] = fmax(v, rb_floor);
    });
    T.sync();
}

// interior-point path (ipm.cuh)
struct IpmOut {
    bool solved, almost, infeasible, blowup, handoff;
    int iters, nfact;
    double rp, rd, rho_p;
};
template <bool RING, class Team>
__device__ IpmOut ipm_run(Team& T, const Inst& I, const CholDev& C, const CholWork& W, Ring* R, const sqpqp_options& o,
                          double c, int phase, const double* xk_scaled_start, int handoff_k = 0, IpmState* st = nullptr,
                          bool resume = false);
__device__ __forceinline__ void ring_init(Ring& R, const CholDev& C, const CholWork& W, int stage0_dbl, unsigned long long* bars);

// ---------------------------------------------------------------------------------------
// MODE 0: interior point, then the ADMM fallback in the same kernel (cooperative-grid team).
// MODE 1: interior point only; an instance it cannot finish is flagged in P.fb_flag for the MODE 2 launch that follows.
// MODE 2: ADMM only, for the flagged instances (or all of them when options.method == 1).
// CTA teams use 1 + 2: the interior-point kernel then carries no ADMM / polish state, which is what keeps the
// 64-register build from spilling (the spill traffic of the monolithic kernel was the larger part of its DRAM traffic,
// profiles/r01_ncu_summary_v2.md).  The work-vector aliases below are scoped per stage for the same reason.
#define SQPQP_ALIASES \
    double *q = I.nv[N_Q], *xl = I.nv[N_XL], *xu = I.nv[N_XU], *D = I.nv[N_D], *x = I.nv[N_X], *zb = I.nv[N_ZB], \
           *yb = I.nv[N_YB], *rb = I.nv[N_RB], *xt = I.nv[N_XT], *rv = I.nv[N_R], *dsh = I.nv[N_DSH], \
           *hd = I.nv[N_HD], *tmpN = I.nv[N_TMP], *tmpN2 = I.nv[N_TMP2], *xw = I.nv[N_XW], *ybw = I.nv[N_YBW]; \
    double *rl = I.mv[M_RL], *ru = I.mv[M_RU], *Es = I.mv[M_ES], *zc = I.mv[M_ZC], *yc = I.mv[M_YC], *rc = I.mv[M_RC], \
           *tmpM = I.mv[M_TMP], *Ax = I.mv[M_AX], *ycw = I.mv[M_YCW]; \
    (void)q; (void)xl; (void)xu; (void)D; (void)x; (void)zb; (void)yb; (void)rb; (void)xt; (void)rv; (void)dsh; (void)hd; \
    (void)tmpN; (void)tmpN2; (void)xw; (void)ybw; (void)rl; (void)ru; (void)Es; (void)zc; (void)yc; (void)rc; (void)tmpM; \
    (void)Ax; (void)ycw;

// RING: the launch may stream its index programs through the shared-memory ring (one resident CTA per SM only: the code is
// compiled out of the other variants, where its registers would turn into spills)
template <int MODE, class Team, bool RING = false>
__device__ void solve_instance(Team& T, const Prob& P, const sqpqp_options& o, int inst, int phase,
                               const Placement* pl, double* dsm, unsigned long long* rbar = nullptr, int handoff_k = 0,
                               bool resume = false) {
    if constexpr (MODE == 2) {
        if (o.method != 1 && P.fb_flag[inst] == 0) return;  // uniform per team: solved by the interior-point launch
    }
    if constexpr (MODE == 1) {
        if (resume && P.fb_flag[inst] != 3) return;         // the resident launch only continues what was handed over
    }
    Prof pfo;
    pfo.start();
    Inst I;
    I.n = P.n; I.m = P.m; I.S = P.S;
    I.M = P.m;
    I.N = (phase == SQPQP_PHASE_FR) ? P.Ne : P.n;
    I.useH = (phase == SQPQP_PHASE_QP || phase == SQPQP_PHASE_SOC) && P.has_hess;
    I.pconst = (phase == SQPQP_PHASE_LP) ? 2.0 : 0.0;
    I.J = Csr{P.J_rb, (phase == SQPQP_PHASE_FR) ? P.J_re_e : P.J_re_n, P.J_col};
    I.T = Csr{P.T_rb, P.T_rb + 1, P.T_col};
    I.H = Csr{P.H_rb, P.H_rb + 1, P.H_col};
    I.lgJ = (phase == SQPQP_PHASE_FR) ? P.lgJe : P.lgJn;
    I.lgT = P.lgT > P.lgH ? P.lgT : P.lgH;
    I.lgH = P.lgH;
    I.Jv = P.Jv + (size_t)inst * P.nnzJ; I.Tv = P.Tv + (size_t)inst * P.nnzT; I.Hv = P.Hv + (size_t)inst * P.nnzH;
    I.Jsv = P.Jsv + (size_t)inst * P.nnzJ; I.Tsv = P.Tsv + (size_t)inst * P.nnzT; I.Hsv = P.Hsv + (size_t)inst * P.nnzH;
    const bool resident = pl && pl->vec_resident;  // scratch arrays resident in this CTA's shared memory for the whole solve
    I.nv = VecTab{P.nv, resident ? pl->n_off : (const int*)nullptr, dsm, (size_t)inst * P.Ne};
    I.mv = VecTab{P.mv, resident ? pl->m_off : (const int*)nullptr, dsm, (size_t)inst * P.m};
    if (resident) {
        if (pl->jsv >= 0) I.Jsv = dsm + pl->jsv;
        if (pl->tsv >= 0) I.Tsv = dsm + pl->tsv;
        if (pl->hsv >= 0) I.Hsv = dsm + pl->hsv;
    }
    I.codeC = P.codeC + (size_t)inst * P.m; I.prevC = P.prevC + (size_t)inst * P.m; I.triedC = P.triedC + (size_t)inst * P.m;
    I.codeB = P.codeB + (size_t)inst * P.Ne; I.prevB = P.prevB + (size_t)inst * P.Ne; I.triedB = P.triedB + (size_t)inst * P.Ne;
    const int n = P.n, m = P.m, N = I.N, M = I.M;
    const double* df = P.df + (size_t)inst * n;
    const double* Ecur = (phase == SQPQP_PHASE_SOC && P.Eov) ? P.Eov + (size_t)inst * m : P.E + (size_t)inst * m;
    const double* gL = P.gL + (size_t)inst * P.gstride;
    const double* gU = P.gU + (size_t)inst * P.gstride;
    const double* xL = P.xL + (size_t)inst * P.xstride;
    const double* xU = P.xU + (size_t)inst * P.xstride;
    const double* xk = P.xk + (size_t)inst * n;
    const double delta = P.delta[inst];

    double c = 1.0;
    {   // ================= stage A: QP data, equilibration (its aliases die with the block) =================
    SQPQP_ALIASES
    // ---- 0. assemble the unscaled QP (set_trust_region!, modify_constraints!) ----------
    for_n(T, N, [&](int j) {
        double lo, hi, qq;
        if (j < n) {
            if (phase == SQPQP_PHASE_LP) {
                lo = xL[j]; hi = xU[j]; qq = -2.0 * xk[j];
            } else {
                double vl = xL[j] - xk[j], vu = xU[j] - xk[j];
                lo = fmax(-delta, vl); hi = fmin(delta, vu);
                if (lo > hi) {  // x_k outside its bounds (subproblem_JuMP.jl:441-444)
                    lo = fmax(-delta, fmin(0.0, vl));
                    hi = fmin(delta, fmax(0.0, vu));
                }
                qq = (phase == SQPQP_PHASE_FR) ? 0.0 : df[j];
            }
        } else {  // FR slack column: free >= 0 unless its row is already satisfied (:365-380)
            int i = P.slack_row[j - n];
            bool feas = (Ecur[i] >= gL[i]) && (Ecur[i] <= gU[i]);
            lo = 0.0; hi = feas ? 0.0 : INFINITY; qq = 1.0;
        }
        xl[j] = lo; xu[j] = hi; q[j] = qq;
    });
    for_n(T, M, [&](int i) {
        double lo, hi;
        if (phase == SQPQP_PHASE_LP) {
            if (i < P.mlin) { lo = gL[i]; hi = gU[i]; } else { lo = -INFINITY; hi = INFINITY; }
        } else {
            lo = gL[i] - Ecur[i]; hi = gU[i] - Ecur[i];
        }
        rl[i] = lo; ru[i] = hi;
    });
    T.sync();

    // ---- 1. Ruiz equilibration (D cols, Es rows, c cost) --------------------------------
    for_n(T, N, [&](int j) { D[j] = 1.0; });
    for_n(T, M, [&](int i) { Es[i] = 1.0; });
    T.sync();
    for (int it = 0; it < o.ruiz_iters; ++it) {
        // column norms (rows of the symmetric P and of the transpose)
        for (int j = T.tid(); j < N; j += T.size()) {
            double a = 0.0;
            if (I.useH)
                for (int k = I.H.rb[j]; k < I.H.re[j]; ++k) a = fmax(a, fabs(I.Hv[k]) * D[I.H.col[k]]);
            a = c * (a + 0.0);
            if (I.pconst != 0.0) a = fmax(a, c * I.pconst * D[j]);
            double b = 0.0;
            for (int k = I.T.rb[j]; k < I.T.re[j]; ++k) b = fmax(b, fabs(I.Tv[k]) * Es[I.T.col[k]]);
            double cn = D[j] * fmax(a, b);
            tmpN[j] = 1.0 / sqrt((cn > 1e-4) ? fmin(cn, 1e4) : 1.0);
        }
        for (int i = T.tid(); i < M; i += T.size()) {
            double a = 0.0;
            for (int k = I.J.rb[i]; k < I.J.re[i]; ++k) a = fmax(a, fabs(I.Jv[k]) * D[I.J.col[k]]);
            double rn = Es[i] * a;
            tmpM[i] = 1.0 / sqrt((rn > 1e-4) ? fmin(rn, 1e4) : 1.0);
        }
        T.sync();
        for_n(T, N, [&](int j) { D[j] *= tmpN[j]; });
        for_n(T, M, [&](int i) { Es[i] *= tmpM[i]; });
        T.sync();
        // cost scaling: mean column norm of P and |q|_inf
        double s[1] = {0.0}, mx[1] = {0.0};
        for (int j = T.tid(); j < N; j += T.size()) {
            double a = 0.0;
            if (I.useH)
                for (int k = I.H.rb[j]; k < I.H.re[j]; ++k) a = fmax(a, fabs(I.Hv[k]) * D[I.H.col[k]]);
            if (I.pconst != 0.0) a = fmax(a, I.pconst * D[j]);
            s[0] += c * D[j] * a;
            mx[0] = fmax(mx[0], fabs(c * D[j] * q[j]));
        }
        T.template reduce<1, false>(s);
        T.template reduce<1, true>(mx);
        double g = fmax(s[0] / (double)N, mx[0]);
        g = 1.0 / ((g > 1e-4) ? g : 1.0);
        g = fmin(fmax(g, 1e-4), 1e4);
        c *= g;
    }
    // scaled values + scaled vectors
    for (int i = T.tid(); i < M; i += T.size())
        for (int k = I.J.rb[i]; k < I.J.re[i]; ++k) I.Jsv[k] = Es[i] * I.Jv[k] * D[I.J.col[k]];
    for (int j = T.tid(); j < N; j += T.size()) {
        for (int k = I.T.rb[j]; k < I.T.re[j]; ++k) I.Tsv[k] = D[j] * I.Tv[k] * Es[I.T.col[k]];
        double dg = c * I.pconst * D[j] * D[j];
        if (I.useH)
            for (int k = I.H.rb[j]; k < I.H.re[j]; ++k) {
                double v = c * D[j] * I.Hv[k] * D[I.H.col[k]];
                I.Hsv[k] = v;
                if (I.H.col[k] == j) dg += v;
            }
        hd[j] = dg;  // diag(Ps) (for !useH the constant diagonal is applied through hd in apply_K)
        q[j] *= c * D[j];
        xl[j] /= D[j];
        xu[j] /= D[j];
    }
    for_n(T, M, [&](int i) { rl[i] *= Es[i]; ru[i] *= Es[i]; });
    T.sync();
    }   // stage A

    int status = SQPQP_MOI_ITERATION_LIMIT;
    int k = 0, cg_total = 0, polish_tries = 0, polish_cg = 0, rho_updates = 0, checks = 0, bumps = 0;
    int ipm_iters = 0, nfact = 0;
    bool polished = false;
    double rp = INFINITY, rd = INFINITY, last_rel = INFINITY;
    double rho = o.rho0, rb_floor = 0.0;

    // ---- 1b. interior point method (default where the symbolic Cholesky is available) ---------
    bool ipm_done = false, ipm_blowup = false;
    const bool fr = phase == SQPQP_PHASE_FR;
    if constexpr (MODE == 2) {  // statistics of the interior-point launch that flagged this instance
        ipm_iters = P.o_info[inst].ipm_iters;
        nfact = P.o_info[inst].chol_factorizations;
        ipm_blowup = P.fb_flag[inst] == 2;
    }
    if (MODE != 2 && (fr ? P.has_chol_fr : P.has_chol) && o.method != 1) {
        const CholDev& CD = fr ? P.chol_fr : P.chol;
        CholWork W;
        W.L = fr ? P.Lval_fr + (size_t)inst * CD.nnzL : P.Lval + (size_t)inst * CD.nnzL;
        W.yw = fr ? P.yw_fr + (size_t)inst * P.Ne : P.yw + (size_t)inst * P.n;
        W.dinv = fr ? P.dinv_fr + (size_t)inst * P.Ne : P.dinv + (size_t)inst * P.n;
        W.wJ = P.wJ + (size_t)inst * P.nnzJ;
        W.D = W.col = W.gsm = nullptr;
        if (!pl) {  // grid team: the dense tail lives in global memory, dsm is the per-CTA scratch of its blocked code
            W.D = fr ? P.Dtail_fr : P.Dtail;
            W.gsm = dsm;
        }
        if (pl) {  // shared-memory parts of the factorisation (the dense tail lives nowhere else)
            if (pl->lval >= 0) W.L = dsm + pl->lval;
            if (pl->yw >= 0) W.yw = dsm + pl->yw;
            if (pl->dinv >= 0) W.dinv = dsm + pl->dinv;
            if (pl->dtail >= 0) { W.D = dsm + pl->dtail; W.col = dsm + pl->dcol; }
        }
        const double* start = nullptr;
        if (phase == SQPQP_PHASE_LP && !resume) {
            double *x = I.nv[N_X], *D = I.nv[N_D];
            for_n(T, N, [&](int j) { x[j] = xk[j] / D[j]; });
            T.sync();
            start = x;
        }
        Ring R;
        R.on = false;
        W.oL = W.oyw = W.odinv = W.oD = -1;
        if constexpr (RING) {
            if (pl && rbar && pl->ring >= 0 && CD.ring_ok) {  // ring mode: L, yw, dinv (and the tail) are in shared memory (launch_solve)
                W.oL = pl->lval; W.oyw = pl->yw; W.odinv = pl->dinv; W.oD = pl->dtail >= 0 ? pl->dtail : 0;
                ring_init(R, CD, W, pl->ring, rbar);
            }
        }
        pfo.lap(PS_PROLOGUE);
        IpmOut io = ipm_run<RING>(T, I, CD, W, (RING && R.on) ? &R : (Ring*)nullptr, o, c, phase, start, handoff_k,
                                  (handoff_k > 0 || resume) ? P.ipm_state + inst : (IpmState*)nullptr, resume);
        pfo.start();
        if (io.handoff) {  // quota used up: state saved, the resident launch continues this instance
            if (T.tid() == 0) P.fb_flag[inst] = 3;
            return;
        }
        ipm_iters = io.iters;
        nfact = io.nfact;
        ipm_blowup = io.blowup;
        if (io.infeasible) {
            ipm_done = true;
            status = SQPQP_MOI_LOCALLY_INFEASIBLE;
        } else if (io.solved || io.almost) {
            ipm_done = true;
            status = io.solved ? SQPQP_MOI_LOCALLY_SOLVED : SQPQP_MOI_ALMOST_LOCALLY_SOLVED;
            rp = io.rp;
            rd = io.rd;
            rb_floor = io.rho_p;
        }
    }
    if constexpr (MODE == 1) {
        if (T.tid() == 0) P.fb_flag[inst] = (!ipm_done && o.method != 2) ? (ipm_blowup ? 2 : 1) : 0;
        if (!ipm_done && o.method != 2) {  // handed to the ADMM launch; keep the interior-point statistics for it
            if (T.tid() == 0) {
                P.o_info[inst].ipm_iters = ipm_iters;
                P.o_info[inst].chol_factorizations = nfact;
            }
            return;
        }
    }
    if (MODE != 1 && !ipm_done && o.method != 2) {
    SQPQP_ALIASES
    // ---- 2. nonconvexity guard: lambda_min(Ps) by power iteration on (bound I - Ps) -------
    if (I.useH) {
        double bnd[1] = {0.0};
        for (int j = T.tid(); j < N; j += T.size()) {
            double a = 0.0;
            for (int k = I.H.rb[j]; k < I.H.re[j]; ++k) a += fabs(I.Hsv[k]);
            bnd[0] = fmax(bnd[0], a);
        }
        T.template reduce<1, true>(bnd);
        if (bnd[0] > 0.0) {
            double* v = tmpN; double* w = tmpN2;
            double nn[1] = {0.0};
            for_n(T, N, [&](int j) { double t = cos(0.7 * (double)j + 0.3); v[j] = t; nn[0] = fma(t, t, nn[0]); });
            T.template reduce<1, false>(nn);
            double inv = 1.0 / sqrt(nn[0]);
            for_n(T, N, [&](int j) { v[j] *= inv; });
            double lam = 0.0;
            for (int it = 0; it < o.eig_iters; ++it) {
                T.sync();
                double rq[2] = {0.0, 0.0};
                csr_rows(T, N, I.lgH, I.H.rb, I.H.re, I.H.col, I.Hsv, v, [&](int r, double d) {
                    double wj = bnd[0] * v[r] - d;
                    w[r] = wj;
                    rq[0] = fma(v[r], wj, rq[0]);
                    rq[1] = fma(wj, wj, rq[1]);
                });
                T.template reduce<2, false>(rq);
                lam = rq[0];
                if (!(rq[1] > 0.0)) break;
                double iw = 1.0 / sqrt(rq[1]);
                for_n(T, N, [&](int j) { v[j] = w[j] * iw; });
            }
            double lmin = bnd[0] - lam;
            if (lmin < 0.0) rb_floor = o.rb_full_mult * (-lmin);
        }
        T.sync();
    }

    // ---- 3. ADMM -------------------------------------------------------------------------
    const double sigma = o.sigma, alpha = o.alpha;
    bool warm = o.warm_start && P.rho_w[inst] > 0.0 && phase != SQPQP_PHASE_LP && phase != SQPQP_PHASE_FR;
    // rho is NOT carried over: a step size adapted to the previous QP (often << rho0) can stall the next one
    if (phase == SQPQP_PHASE_LP) {
        for_n(T, N, [&](int j) { x[j] = xk[j] / D[j]; yb[j] = 0.0; });
        for_n(T, M, [&](int i) { yc[i] = 0.0; });
    } else if (warm) {
        for_n(T, N, [&](int j) { x[j] = xw[j] / D[j]; yb[j] = c * ybw[j] * D[j]; });
        for_n(T, M, [&](int i) { yc[i] = c * ycw[i] / Es[i]; });
    } else {
        for_n(T, N, [&](int j) { x[j] = 0.0; yb[j] = 0.0; });
        for_n(T, M, [&](int i) { yc[i] = 0.0; });
    }
    T.sync();
    csr_rows(T, M, I.lgJ, I.J.rb, I.J.re, I.J.col, I.Jsv, x, [&](int r, double d) { zc[r] = clampd(d, rl[r], ru[r]); });
    for_n(T, N, [&](int j) { zb[j] = clampd(x[j], xl[j], xu[j]); xt[j] = x[j]; });
    for_n(T, M, [&](int i) { I.prevC[i] = -1; I.triedC[i] = -1; });
    for_n(T, N, [&](int j) { I.prevB[j] = -1; I.triedB[j] = -1; });
    set_rho(T, I, o, rho, rb_floor);
    for_n(T, N, [&](int j) { dsh[j] = sigma + rb[j]; });
    T.sync();
    build_minv(T, I, dsh, rc);

    double* dyc = I.mv[M_BC];   // dy of the last iteration (rows); M_BC is only used by polish afterwards
    double* dyb = I.nv[N_XFIX]; // dy (box)
    const int admm_cap = (nfact > 0 && !ipm_blowup && o.fallback_max_iter > 0 && o.fallback_max_iter < o.max_iter) ? o.fallback_max_iter : o.max_iter;
    while (k < admm_cap) {
        ++k;
        // rhs = sigma x - q + Ts (rc zc - yc) + (rb zb - yb)
        double* t = I.mv[M_T];
        for_n(T, M, [&](int i) { t[i] = rc[i] * zc[i] - yc[i]; });
        T.sync();
        double nr[1] = {0.0};
        csr_rows(T, N, I.lgT, I.T.rb, I.T.re, I.T.col, I.Tsv, t, [&](int r, double d) {
            double b = sigma * x[r] - q[r] + d + rb[r] * zb[r] - yb[r];
            rv[r] = b;
            nr[0] = fma(b, b, nr[0]);
        });
        T.template reduce<1, false>(nr);
        CgOut cgo = pcg(T, I, xt, dsh, rc, (const double*)nullptr, fmax(1e-14 * sqrt(nr[0]), 1e-300), o.cg_rel0, o.cg_max);
        cg_total += cgo.iters;
        if (cgo.neg) {  // K not positive definite: raise the box step-size floor and redo
            if (++bumps > 40) { status = SQPQP_MOI_NUMERICAL_ERROR; break; }
            rb_floor = fmax(2.0 * rb_floor, 2.0 * fabs(cgo.curv) + 1e-3);
            set_rho(T, I, o, rho, rb_floor);
            for_n(T, N, [&](int j) { dsh[j] = sigma + rb[j]; xt[j] = x[j]; });
            T.sync();
            build_minv(T, I, dsh, rc);
            --k;
            continue;
        }
        // z~ = Js x~ ; relaxation, projection, dual update
        csr_rows(T, M, I.lgJ, I.J.rb, I.J.re, I.J.col, I.Jsv, xt, [&](int r, double d) {
            double zr = alpha * d + (1.0 - alpha) * zc[r];
            double zn = clampd(zr + yc[r] / rc[r], rl[r], ru[r]);
            double dy = rc[r] * (zr - zn);
            zc[r] = zn;
            yc[r] += dy;
            dyc[r] = dy;
        });
        for_n(T, N, [&](int j) {
            double xtj = xt[j];
            double zr = alpha * xtj + (1.0 - alpha) * zb[j];
            x[j] = alpha * xtj + (1.0 - alpha) * x[j];
            double zn = clampd(zr + yb[j] / rb[j], xl[j], xu[j]);
            double dy = rb[j] * (zr - zn);
            zb[j] = zn;
            yb[j] += dy;
            dyb[j] = dy;
        });
        if (k % o.check_every) { T.sync(); continue; }
        // ---- residual check (unscaled) ----
        ++checks;
        T.sync();
        csr_rows(T, M, I.lgJ, I.J.rb, I.J.re, I.J.col, I.Jsv, x, [&](int r, double d) { Ax[r] = d; });
        double mx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        // [0] rp  [1] rd*c  [2] |Ax|,|z| prim norm  [3] |Px|,|ATy|,|q| dual norm *c  [4] |dy|  [5] |AT dy|
        csr_rows2(T, N, I.lgT, I.H.rb, I.H.re, I.H.col, I.Hsv, x, I.useH, I.T.rb, I.T.re, I.T.col, I.Tsv, yc,
                  [&](int r, double px, double aty) {
                      if (!I.useH) px = hd[r] * x[r];
                      aty += yb[r];
                      double id = 1.0 / D[r];
                      mx[1] = fmax(mx[1], fabs(px + q[r] + aty) * id);
                      mx[3] = fmax(mx[3], fmax(fabs(px), fmax(fabs(aty), fabs(q[r]))) * id);
                      mx[0] = fmax(mx[0], fabs(x[r] - zb[r]) * D[r]);
                      mx[2] = fmax(mx[2], fmax(fabs(x[r]), fabs(zb[r])) * D[r]);
                      mx[4] = fmax(mx[4], fabs(dyb[r]) * id);
                  });
        csr_rows(T, N, I.lgT, I.T.rb, I.T.re, I.T.col, I.Tsv, dyc,
                 [&](int r, double d) { mx[5] = fmax(mx[5], fabs(d + dyb[r]) / D[r]); });
        T.sync();
        for_n(T, M, [&](int i) {
            double ie = 1.0 / Es[i];
            mx[0] = fmax(mx[0], fabs(Ax[i] - zc[i]) * ie);
            mx[2] = fmax(mx[2], fmax(fabs(Ax[i]), fabs(zc[i])) * ie);
            mx[4] = fmax(mx[4], fabs(dyc[i]) * Es[i]);
        });
        T.template reduce<6, true>(*reinterpret_cast<double(*)[6]>(mx));
        rp = mx[0];
        rd = mx[1] / c;
        double np_ = mx[2], nd_ = mx[3] / c;
        double relp = rp / fmax(np_, 1e-30), reld = rd / fmax(nd_, 1e-30);
        last_rel = fmax(relp, reld);
        // ---- primal infeasibility certificate (unscaled dy) ----
        double ndy = mx[4] / c;
        if (ndy > 1e-30) {
            double atdy = mx[5] / c;
            double sp[1] = {0.0};
            for_n(T, M, [&](int i) {
                double d = dyc[i] * Es[i] / c;  // unscaled
                double lo = rl[i] / Es[i], hi = ru[i] / Es[i];
                if (d > 0.0) sp[0] += hi * d; else if (d < 0.0) sp[0] += lo * d;
            });
            for_n(T, N, [&](int j) {
                double d = dyb[j] / (D[j] * c);
                double lo = xl[j] * D[j], hi = xu[j] * D[j];
                if (d > 0.0) sp[0] += hi * d; else if (d < 0.0) sp[0] += lo * d;
            });
            T.template reduce<1, false>(sp);
            if (atdy <= o.eps_inf * ndy && sp[0] <= -o.eps_inf * ndy) {
                status = SQPQP_MOI_LOCALLY_INFEASIBLE;
                break;
            }
        }
        // ---- active-set prediction; polish when it is stable and not yet tried ----
        double cnt[2] = {0.0, 0.0};
        for_n(T, M, [&](int i) {
            signed char cd = 0;
            if (rl[i] == ru[i] || zc[i] - rl[i] < -yc[i]) cd = 1;
            else if (ru[i] - zc[i] < yc[i]) cd = 2;
            cnt[0] += (cd != I.prevC[i]);
            cnt[1] += (cd != I.triedC[i]);
            I.prevC[i] = cd;
            I.codeC[i] = cd;
        });
        for_n(T, N, [&](int j) {
            signed char cd = 0;
            if (xl[j] == xu[j] || x[j] - xl[j] < -yb[j]) cd = 1;
            else if (xu[j] - x[j] < yb[j]) cd = 2;
            cnt[0] += (cd != I.prevB[j]);
            cnt[1] += (cd != I.triedB[j]);
            I.prevB[j] = cd;
            I.codeB[j] = cd;
        });
        T.template reduce<2, false>(cnt);
        bool converged = (rp <= o.eps_abs + o.eps_rel * np_) && (rd <= o.eps_abs + o.eps_rel * nd_);
        if ((fmax(relp, reld) < o.polish_trigger && cnt[0] == 0.0 && cnt[1] != 0.0) || (converged && cnt[1] != 0.0)) {
            ++polish_tries;
            for_n(T, M, [&](int i) { I.triedC[i] = I.codeC[i]; });
            for_n(T, N, [&](int j) { I.triedB[j] = I.codeB[j]; });
            // ---- 4. polish on the predicted active set ----
            double *bc = I.mv[M_BC], *yp = I.mv[M_YP], *rwp = I.mv[M_RW], *xp = I.nv[N_XFIX], *mask = I.nv[N_MASK],
                   *dsp = I.nv[N_TMP], *xfix = I.nv[N_TMP2];
            const double rhoP = o.polish_rho, sigP = 1e-9;
            for_n(T, M, [&](int i) {
                signed char cd = I.codeC[i];
                bc[i] = (cd == 1) ? rl[i] : ((cd == 2) ? ru[i] : 0.0);
                rwp[i] = cd ? rhoP : 0.0;
                yp[i] = cd ? yc[i] : 0.0;
            });
            for_n(T, N, [&](int j) {
                signed char cd = I.codeB[j];
                double f = (cd == 1) ? xl[j] : ((cd == 2) ? xu[j] : 0.0);
                xfix[j] = f;
                mask[j] = cd ? 0.0 : 1.0;
                xp[j] = cd ? f : x[j];
                dsp[j] = sigP;
            });
            T.sync();
            build_minv(T, I, dsp, rwp);
            bool ok = true;
            double resn = INFINITY;
            for (int po = 0; po < o.polish_outer; ++po) {
                // rhs = sigP xp - q + Ts (rwp.*bc - yp)
                double* t2 = I.mv[M_T];
                for_n(T, M, [&](int i) { t2[i] = rwp[i] * bc[i] - yp[i]; });
                T.sync();
                double nb[1] = {0.0};
                csr_rows(T, N, I.lgT, I.T.rb, I.T.re, I.T.col, I.Tsv, t2, [&](int r, double d) {
                    double b = sigP * xp[r] - q[r] + d;
                    rv[r] = b;
                    nb[0] = fma(b, b, nb[0]);
                });
                T.template reduce<1, false>(nb);
                CgOut pc = pcg(T, I, xp, dsp, rwp, mask, fmax(1e-13 * sqrt(nb[0]), 1e-300), 0.0, o.polish_cg_max);
                polish_cg += pc.iters;
                if (pc.neg) { ok = false; break; }
                // xp: free part from CG, fixed part stays (mask keeps it untouched)
                double rr[1] = {0.0};
                csr_rows(T, M, I.lgJ, I.J.rb, I.J.re, I.J.col, I.Jsv, xp, [&](int r, double d) {
                    Ax[r] = d;
                    if (I.codeC[r]) {
                        double res = d - bc[r];
                        yp[r] += rhoP * res;
                        rr[0] = fmax(rr[0], fabs(res));
                    }
                });
                T.template reduce<1, true>(rr);
                resn = rr[0];
                if (resn < o.polish_tol) break;
            }
            if (ok) {
                // verification: stationarity on free cols, primal feasibility, dual signs
                T.sync();
                // [0] stat [1] pf [2] dual-sign [3] |y|max (scaled)  [4] stat [5] pf [6] dual-sign (unscaled*c)
                double vmx[7] = {0, 0, 0, 0, 0, 0, 0};
                double* ybn = I.nv[N_KP];
                csr_rows2(T, N, I.lgT, I.H.rb, I.H.re, I.H.col, I.Hsv, xp, I.useH, I.T.rb, I.T.re, I.T.col, I.Tsv, yp,
                          [&](int r, double px, double aty) {
                              if (!I.useH) px = hd[r] * xp[r];
                              double g = px + q[r] + aty;
                              signed char cd = I.codeB[r];
                              double y = cd ? -g : 0.0;
                              ybn[r] = y;
                              if (!cd) { vmx[0] = fmax(vmx[0], fabs(g)); vmx[4] = fmax(vmx[4], fabs(g) / D[r]); }
                              double pv = fmax(xl[r] - xp[r], xp[r] - xu[r]);
                              vmx[1] = fmax(vmx[1], pv);
                              vmx[5] = fmax(vmx[5], pv * D[r]);
                              bool eqb = xl[r] == xu[r];
                              double dv = 0.0;
                              if (cd == 1 && !eqb) dv = y;
                              if (cd == 2) dv = -y;
                              vmx[2] = fmax(vmx[2], dv);
                              vmx[6] = fmax(vmx[6], dv / D[r]);
                              vmx[3] = fmax(vmx[3], fabs(y));
                          });
                for_n(T, M, [&](int i) {
                    double pv = fmax(rl[i] - Ax[i], Ax[i] - ru[i]);
                    vmx[1] = fmax(vmx[1], pv);
                    vmx[5] = fmax(vmx[5], pv / Es[i]);
                    signed char cd = I.codeC[i];
                    bool eqc = rl[i] == ru[i];
                    double dv = 0.0;
                    if (cd == 1 && !eqc) dv = yp[i];
                    if (cd == 2) dv = -yp[i];
                    vmx[2] = fmax(vmx[2], dv);
                    vmx[6] = fmax(vmx[6], dv * Es[i]);
                    vmx[3] = fmax(vmx[3], fabs(yp[i]));
                });
                T.template reduce<7, true>(vmx);
                double ymag = fmax(1.0, vmx[3]);
                ok = (vmx[1] <= o.feas_tol) && (vmx[2] <= o.dual_tol * ymag) && (resn <= 100.0 * o.polish_tol) &&
                     (vmx[0] <= 1e-9 * ymag);
                // once ADMM itself has converged, a refinement that is no worse than the ADMM point in
                // every unscaled residual is accepted as well (weakly active constraints can fail the
                // strict sign test by rounding)
                if (!ok && converged && resn <= 100.0 * o.polish_tol)
                    ok = (fmax(vmx[5], 0.0) <= rp) && (fmax(vmx[4], vmx[6]) / c <= rd);
                if (ok) {
                    for_n(T, N, [&](int j) { x[j] = xp[j]; yb[j] = ybn[j]; zb[j] = xp[j]; });
                    for_n(T, M, [&](int i) { yc[i] = yp[i]; zc[i] = Ax[i]; });
                    polished = true;
                    status = SQPQP_MOI_LOCALLY_SOLVED;
                    rp = fmax(vmx[5], 0.0);
                    rd = vmx[4] / c;
                    T.sync();
                    break;
                }
            }
            // polish rejected: restore the ADMM preconditioner and continue
            T.sync();
            build_minv(T, I, dsh, rc);
            for_n(T, N, [&](int j) { xt[j] = x[j]; });
            T.sync();
        }
        if (converged) { status = SQPQP_MOI_LOCALLY_SOLVED; break; }
        // ---- rho adaptation ----
        double nrho = rho * sqrt(relp / fmax(reld, 1e-30));
        nrho = fmin(fmax(nrho, o.rho_min), o.rho_max);
        if (nrho > rho * o.adapt_tol || nrho < rho / o.adapt_tol) {
            rho = nrho;
            ++rho_updates;
            set_rho(T, I, o, rho, rb_floor);
            for_n(T, N, [&](int j) { dsh[j] = sigma + rb[j]; });
            T.sync();
            build_minv(T, I, dsh, rc);
        }
        T.sync();
    }
    T.sync();
    // ALMOST_LOCALLY_SOLVED means what it means in the interior-point path and in Ipopt ("acceptable level"): both
    // unscaled residuals at 1e-6 relative.  Anything looser stays ITERATION_LIMIT, which the SQP driver treats as an
    // unexpected sub-status (sqp_trust_region.jl:169-177) instead of consuming the step and its multipliers.
    if (status == SQPQP_MOI_ITERATION_LIMIT && last_rel <= 1e-6) status = SQPQP_MOI_ALMOST_LOCALLY_SOLVED;
    }  // ADMM path
    else if (!ipm_done) status = SQPQP_MOI_NUMERICAL_ERROR;

    // ---- 5. outputs (collect_solution!, subproblem_JuMP.jl:514-563) -----------------------
    SQPQP_ALIASES
    bool okst = (status == SQPQP_MOI_LOCALLY_SOLVED || status == SQPQP_MOI_ALMOST_LOCALLY_SOLVED ||
                 status == SQPQP_MOI_ITERATION_LIMIT);
    double obj[1] = {0.0};
    if (okst) {
        // objective 1/2 x'Px + q'x in unscaled units: (1/c) * scaled objective
        if (I.useH)
            csr_rows(T, N, I.lgH, I.H.rb, I.H.re, I.H.col, I.Hsv, x,
                     [&](int r, double d) { obj[0] += x[r] * (0.5 * d + q[r]); });
        else
            for_n(T, N, [&](int r) { obj[0] += x[r] * (0.5 * hd[r] * x[r] + q[r]); });
        T.template reduce<1, false>(obj);
        obj[0] /= c;
    }
    double* op = P.o_p + (size_t)inst * n;
    double* ol = P.o_lam + (size_t)inst * m;
    double* oL = P.o_mxL + (size_t)inst * n;
    double* oU = P.o_mxU + (size_t)inst * n;
    double* os = P.o_slack + (size_t)inst * (P.S > 0 ? P.S : 1);
    for_n(T, n, [&](int j) {
        double pv = 0.0, rcost = 0.0;
        // a fixed column (lb == ub: trust-region box collapsed onto a bound, or a slack column fixed by the generic lane)
        // returns its bound exactly, as a solver that eliminates fixed variables does
        if (okst) { pv = D[j] * ((xl[j] == xu[j]) ? xl[j] : x[j]); rcost = -yb[j] / (D[j] * c); }
        op[j] = pv;
        oL[j] = rcost > 0.0 ? rcost : 0.0;
        oU[j] = rcost < 0.0 ? rcost : 0.0;
    });
    for_n(T, m, [&](int i) { ol[i] = okst ? -(Es[i] * yc[i]) / c : 0.0; });
    for_n(T, P.S, [&](int s) { os[s] = (okst && phase == SQPQP_PHASE_FR) ? D[n + s] * ((xl[n + s] == xu[n + s]) ? xl[n + s] : x[n + s]) : 0.0; });
    // warm start for the next solve of this instance (unscaled)
    if (okst && phase != SQPQP_PHASE_LP && phase != SQPQP_PHASE_FR) {
        for_n(T, n, [&](int j) { xw[j] = D[j] * x[j]; ybw[j] = yb[j] / (D[j] * c); });
        for_n(T, m, [&](int i) { ycw[i] = Es[i] * yc[i] / c; });
    }
    if (T.tid() == 0) {
        if (okst && phase != SQPQP_PHASE_LP && phase != SQPQP_PHASE_FR) P.rho_w[inst] = rho;
        sqpqp_info& inf = P.o_info[inst];
        inf.moi_status = status;
        inf.admm_iters = k;
        inf.cg_iters = cg_total;
        inf.polish_tries = polish_tries;
        inf.polish_cg_iters = polish_cg;
        inf.polished = polished ? 1 : 0;
        inf.rho_updates = rho_updates;
        inf.checks = checks;
        inf.rho = rho;
        inf.rho_box_floor = rb_floor;
        inf.res_prim = rp;
        inf.res_dual = rd;
        inf.objective = obj[0];
        inf.ipm_iters = ipm_iters;
        inf.chol_factorizations = nfact;
    }
    T.sync();
    pfo.lap(PS_EPILOGUE);
}

// ---- kernels ---------------------------------------------------------------------------
// MAXT threads per CTA, at least MINB CTAs per SM: <512,1> keeps 128 registers/thread for the
// one-CTA-per-SM resident configuration; <256,4> and <128,8> cap registers at 64 so that many
// instances share an SM and hide each other's (L2-latency-bound) level-scheduled phases.
template <int MAXT, int MINB, int MODE>
__global__ void __launch_bounds__(MAXT, MINB) k_solve_cta(const __grid_constant__ Prob P, const __grid_constant__ DevOpts O, int phase,
                                                          const __grid_constant__ Placement pl) {
    __shared__ double sh[2 * SQPQP_MAX_RED * 32];
    __shared__ unsigned long long rbar[RING_S];  // mbarriers of the index-program ring (chol.cuh)
    extern __shared__ double dsm[];
    for (int slot = blockIdx.x; slot < P.batch; slot += gridDim.x) {
        const int inst = O.order ? O.order[slot] : slot;
        if (P.active && !P.active[inst]) continue;
        CtaTeam T(sh);
        solve_instance<MODE, CtaTeam, (MINB == 1 && MODE == 1)>(T, P, O.o, inst, phase, &pl, dsm, rbar, MODE == 1 ? O.handoff_k : 0,
                                                                MODE == 1 && O.resume != 0);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) k_solve_grid(const __grid_constant__ Prob P, const __grid_constant__ DevOpts O, int phase) {
    __shared__ double sh[SQPQP_MAX_RED * 32];
    __shared__ double gsm[2 * GD_NB * GD_LD + 64];  // blocked dense tail: two 32 x 33 tiles, inverse pivots, a right-hand-side block
    GridTeam T(sh, P.gred, P.gred_stride);
    for (int inst = 0; inst < P.batch; ++inst) {
        if (P.active && !P.active[inst]) continue;
        solve_instance<0>(T, P, O.o, inst, phase, (const Placement*)nullptr, gsm);
    }
}
