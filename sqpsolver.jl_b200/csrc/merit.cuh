// merit.cuh -- the arithmetic around each QP solve that the SQP loop needs (rows M1-M3 of
// SURVEY.md section 8a), on device, one CTA per instance, deterministic reductions.
//
//   norm_violations     common.jl:54-77  (p = 1)
//   compute_phi         sqp.jl:170-183   (f, g at the trial point come from the host callbacks)
//   compute_qmodel      sqp_trust_region.jl:487-508
//   KT_residuals        common.jl:14-23  (as coded, including the sign convention)
#pragma once
#include "team.cuh"

__device__ __forceinline__ double viol1(double v, double lo, double hi) {
    return (v > hi) ? (v - hi) : ((v < lo) ? (lo - v) : 0.0);
}

struct MeritArgs {
    const double *x, *p, *Etrial, *ftrial, *mu;
    const int* fr;
    double *viol0, *violt, *phit, *q0, *qk;
};

__global__ void __launch_bounds__(256) k_merit(Prob P, MeritArgs A) {
    __shared__ double sh[2 * SQPQP_MAX_RED * 32];
    for (int inst = blockIdx.x; inst < P.batch; inst += gridDim.x) {
        if (P.active && !P.active[inst]) continue;
        CtaTeam T(sh);
        const int n = P.n, m = P.m;
        const double* x = A.x + (size_t)inst * n;
        const double* p = A.p + (size_t)inst * n;
        const double* Et = A.Etrial + (size_t)inst * m;
        const double* E = P.E + (size_t)inst * m;
        const double* df = P.df + (size_t)inst * n;
        const double* gL = P.gL + (size_t)inst * P.gstride;
        const double* gU = P.gU + (size_t)inst * P.gstride;
        const double* xL = P.xL + (size_t)inst * P.xstride;
        const double* xU = P.xU + (size_t)inst * P.xstride;
        const double* Jv = P.Jv + (size_t)inst * P.nnzJ;
        const double* Hv = P.Hv + (size_t)inst * P.nnzH;
        // s[0]=viol(E,x) s[1]=viol(Etrial,x+p) s[2]=viol(E+Jp,x+p) s[3]=df'p s[4]=p'Hp
        double s[5] = {0, 0, 0, 0, 0};
        csr_rows(T, m, P.lgJn, P.J_rb, P.J_re_n, P.J_col, Jv, p, [&](int i, double jp) {
            s[0] += viol1(E[i], gL[i], gU[i]);
            s[1] += viol1(Et[i], gL[i], gU[i]);
            s[2] += viol1(E[i] + jp, gL[i], gU[i]);
        });
        if (P.has_hess)
            csr_rows(T, n, P.lgH, P.H_rb, P.H_rb + 1, P.H_col, Hv, p, [&](int j, double hp) { s[4] = fma(p[j], hp, s[4]); });
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            double xt = x[j] + p[j];
            double vx = viol1(xt, xL[j], xU[j]);
            s[0] += viol1(x[j], xL[j], xU[j]);
            s[1] += vx;
            s[2] += vx;
            s[3] = fma(df[j], p[j], s[3]);
        }
        T.reduce<5, false>(s);
        if (threadIdx.x == 0) {
            double mu = A.mu[inst];
            bool fr = A.fr && A.fr[inst];
            if (A.viol0) A.viol0[inst] = s[0];
            if (A.violt) A.violt[inst] = s[1];
            if (A.phit) A.phit[inst] = fr ? s[1] : A.ftrial[inst] + mu * s[1];
            if (A.q0) A.q0[inst] = mu * s[0];
            if (A.qk) A.qk[inst] = s[3] + 0.5 * s[4] + mu * s[2];
        }
        __syncthreads();
    }
}

// Line-search primitives (the reference's line-search driver is not compiled -- sqp.jl:226 -- so this is the set of
// device quantities its merit maths needs: compute_mu_rule2! sqp_line_search.jl:280-291, compute_alpha :303-334,
// compute_phi sqp.jl:170-183, compute_derivative sqp.jl:190-213 + merit.jl:13-17, norm_complementarity common.jl:30-47).
struct LsArgs {
    const double *x, *p, *alpha, *Etrial, *mu, *lam;
    double* out;  // [8][batch]: df'p, p'Hp, |viol(E,x)|_1, |viol(E,x)|_inf, weighted viol at x, weighted viol at the trial
                  //             point, |viol(E_trial, x + alpha p)|_1, normalised complementarity (p = Inf)
};

__global__ void __launch_bounds__(256) k_linesearch(Prob P, LsArgs A) {
    __shared__ double sh[2 * SQPQP_MAX_RED * 32];
    for (int inst = blockIdx.x; inst < P.batch; inst += gridDim.x) {
        CtaTeam T(sh);
        const int n = P.n, m = P.m;
        const double* x = A.x + (size_t)inst * n;
        const double* p = A.p + (size_t)inst * n;
        const double al = A.alpha[inst];
        const double* Et = A.Etrial + (size_t)inst * m;
        const double* mu = A.mu + (size_t)inst * m;
        const double* lam = A.lam + (size_t)inst * m;
        const double* E = P.E + (size_t)inst * m;
        const double* df = P.df + (size_t)inst * n;
        const double* gL = P.gL + (size_t)inst * P.gstride;
        const double* gU = P.gU + (size_t)inst * P.gstride;
        const double* xL = P.xL + (size_t)inst * P.xstride;
        const double* xU = P.xU + (size_t)inst * P.xstride;
        const double* Hv = P.Hv + (size_t)inst * P.nnzH;
        // sums: [0] df'p [1] p'Hp [2] row viol at E [3] mu'viol(E) [4] mu'viol(E_trial) [5] row viol at E_trial
        //       [6] bound viol at x [7] bound viol at x + alpha p
        double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        // maxima: [0] |viol(E,x)|_inf [1] |mu|_inf [2] |compl|_inf ; sum: lambda^2 over inequality rows
        double mx[3] = {0, 0, 0}, l2[1] = {0};
        if (P.has_hess)
            csr_rows(T, n, P.lgH, P.H_rb, P.H_rb + 1, P.H_col, Hv, p, [&](int j, double hp) { s[1] = fma(p[j], hp, s[1]); });
        for (int i = threadIdx.x; i < m; i += blockDim.x) {
            const double e = E[i], lo = gL[i], hi = gU[i], v0 = viol1(e, lo, hi), vt = viol1(Et[i], lo, hi), mi = mu[i];
            s[2] += v0; s[3] = fma(mi, v0, s[3]); s[4] = fma(mi, vt, s[4]); s[5] += vt;
            mx[0] = fmax(mx[0], v0);
            mx[1] = fmax(mx[1], fabs(mi));
            if (lo != hi) {
                const double cp = fmin(e - lo, hi - e) * lam[i];
                if (cp == cp) mx[2] = fmax(mx[2], fabs(cp));  // inf * 0 on a one-sided row with a zero multiplier
                l2[0] = fma(lam[i], lam[i], l2[0]);
            }
        }
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            const double v0 = viol1(x[j], xL[j], xU[j]), vt = viol1(fma(al, p[j], x[j]), xL[j], xU[j]);
            s[6] += v0; s[7] += vt;
            mx[0] = fmax(mx[0], v0);
            s[0] = fma(df[j], p[j], s[0]);
        }
        T.reduce<8, false>(s);
        T.reduce<3, true>(mx);
        T.reduce<1, false>(l2);
        if (threadIdx.x == 0) {
            const size_t B = P.batch;
            double* o = A.out + inst;
            o[0] = s[0]; o[B] = s[1]; o[2 * B] = s[2] + s[6]; o[3 * B] = mx[0];
            o[4 * B] = s[3] + mx[1] * s[6]; o[5 * B] = s[4] + mx[1] * s[7]; o[6 * B] = s[5] + s[7];
            o[7 * B] = mx[2] / (1.0 + sqrt(l2[0]));
        }
        __syncthreads();
    }
}

struct KtArgs {
    const double *lam, *mxU, *mxL;
    double* kt;
};

__global__ void __launch_bounds__(256) k_kt(Prob P, KtArgs A) {
    __shared__ double sh[2 * SQPQP_MAX_RED * 32];
    for (int inst = blockIdx.x; inst < P.batch; inst += gridDim.x) {
        if (P.active && !P.active[inst]) continue;
        CtaTeam T(sh);
        const int n = P.n, m = P.m;
        const double* lam = A.lam + (size_t)inst * m;
        const double* mxU = A.mxU + (size_t)inst * n;
        const double* mxL = A.mxL + (size_t)inst * n;
        const double* df = P.df + (size_t)inst * n;
        const double* Jv = P.Jv + (size_t)inst * P.nnzJ;
        const double* Tv = P.Tv + (size_t)inst * P.nnzT;
        double mx[2] = {0.0, 1.0};  // [0] |df + J'lam + mxU - mxL|_inf   [1] scalar
        csr_rows(T, n, P.lgT, P.T_rb, P.T_rb + 1, P.T_col, Tv, lam, [&](int j, double jtl) {
            mx[0] = fmax(mx[0], fabs(df[j] + jtl + mxU[j] - mxL[j]));
            mx[1] = fmax(mx[1], fmax(fabs(df[j]), fmax(fabs(mxU[j]), fabs(mxL[j]))));
        });
        for (int i = threadIdx.x; i < m; i += blockDim.x) {
            double ss = 0.0;
            for (int k = P.J_rb[i]; k < P.J_re_n[i]; ++k) ss = fma(Jv[k], Jv[k], ss);
            mx[1] = fmax(mx[1], fabs(lam[i]) * sqrt(ss));
        }
        T.reduce<2, true>(mx);
        if (threadIdx.x == 0) A.kt[inst] = mx[0] / mx[1];
        __syncthreads();
    }
}

// J p per instance (``Jacobian * p``, sqp_trust_region.jl:343, 492): out[b][m]
__global__ void __launch_bounds__(256) k_jac_times(Prob P, const double* __restrict__ pv, double* __restrict__ out) {
    for (int inst = blockIdx.x; inst < P.batch; inst += gridDim.x) {
        CtaTeam T(nullptr);
        const double* p = pv + (size_t)inst * P.n;
        const double* Jv = P.Jv + (size_t)inst * P.nnzJ;
        double* o = out + (size_t)inst * P.m;
        csr_rows(T, P.m, P.lgJn, P.J_rb, P.J_re_n, P.J_col, Jv, p, [&](int i, double d) { o[i] = d; });
    }
}
