// TEST INFRASTRUCTURE ONLY (not part of libsqpqp.so, never on the product path).
// Executes the index programs produced by csrc/symbolic.hpp on the host, in the order the
// device kernels execute them (csrc/chol.cuh: assembly, level-scheduled sparse columns, Schur
// complement of the dense tail, dense packed Cholesky of the tail, forward / tail / backward
// sweeps), so that tests/test_symbolic.py can validate the symbolic analysis against a dense
// Cholesky on a machine without a GPU.
#include <cmath>
#include <cstdint>
#include <vector>
#include "../../sqpsolver.jl_b200/csrc/symbolic.hpp"

static inline int tri(int r) { return r * (r + 1) / 2; }

extern "C" int symcheck_solve2(int n, int m, const int* Jrp, const int* Jcol, const double* Jv, const int* Prp, const int* Pcol,
                               const double* Pv, const double* d, const double* w, const double* rhs, double* x,
                               int64_t* stats /* nnzL, sparse levels, flops, assembly terms, tail, tree levels */, int tail_max) {
    Symbolic S = symbolic_analyze(n, m, Jrp, Jrp + 1, Jcol, Prp, Pcol, 512, tail_max);
    if (!S.ok) return -1;
    const int n0 = S.n0, T = S.T;
    std::vector<double> L(S.nnzL), D((size_t)tri(T) + T, 0.0), dinv(n);
    for (int e = 0; e < S.nnzL; ++e) {
        double v = 0.0;
        const int* hd = &S.as_hd[4 * (size_t)e];
        if (hd[0] >= 0) v += Pv[hd[0]];
        if (hd[1] >= 0) v += d[hd[1]];
        for (int t = hd[2]; t < hd[3]; ++t) {
            const int* abr = &S.as_abr[4 * (size_t)t];
            v += w[abr[2]] * Jv[abr[0]] * Jv[abr[1]];
        }
        L[e] = v;
    }
    auto pairs = [&](int e) {
        double acc = 0.0;
        for (int q = S.fp_ptr[e]; q < S.fp_ptr[e + 1]; ++q) acc += L[S.fp_ab[2 * (size_t)q]] * L[S.fp_ab[2 * (size_t)q + 1]];
        return acc;
    };
    for (int l = 0; l < S.nlev; ++l)
        for (int j = S.lev_ptr[l]; j < S.lev_ptr[l + 1]; ++j) {
            int e0 = S.Lp[j];
            double dd = L[e0] - pairs(e0);
            if (!(dd > 0.0)) return -2;
            double inv = 1.0 / std::sqrt(dd);
            for (int e = e0 + 1; e < S.Lp[j + 1]; ++e) L[e] = (L[e] - pairs(e)) * inv;
            L[e0] = dd * inv;
            dinv[j] = inv;
        }
    if (S.nlev > 0 && S.lev_ptr[S.nlev] != n0) return -3;
    if (T > 0) {
        const int base = S.Lp[n0];
        for (int e = base; e < S.nnzL; ++e) D[S.tpos[e - base]] = L[e] - pairs(e);
        for (int j = 0; j < T; ++j) {
            double dd = D[tri(j) + j];
            if (!(dd > 0.0)) return -2;
            double inv = 1.0 / std::sqrt(dd);
            dinv[n0 + j] = inv;
            for (int i = j + 1; i < T; ++i) D[tri(i) + j] *= inv;
            for (int i = j + 1; i < T; ++i)
                for (int k = j + 1; k <= i; ++k) D[tri(i) + k] -= D[tri(i) + j] * D[tri(k) + j];
        }
    }
    std::vector<double> y(n);
    for (int k = 0; k < n; ++k) y[k] = rhs[S.perm[k]];
    for (int j = 0; j < n0; ++j) {
        double acc = 0.0;
        for (int q = S.Rp[j]; q < S.Rp[j + 1]; ++q) acc += L[S.Rci[2 * (size_t)q]] * y[S.Rci[2 * (size_t)q + 1]];
        y[j] = (y[j] - acc) * dinv[j];
    }
    if (T > 0) {
        for (int j = n0; j < n; ++j) {
            double acc = 0.0;
            for (int q = S.Rp[j]; q < S.Rmid[j]; ++q) acc += L[S.Rci[2 * (size_t)q]] * y[S.Rci[2 * (size_t)q + 1]];
            y[j] -= acc;
        }
        double* t = &y[n0];
        for (int j = 0; j < T; ++j) {
            t[j] *= dinv[n0 + j];
            for (int i = j + 1; i < T; ++i) t[i] -= D[tri(i) + j] * t[j];
        }
        for (int j = T - 1; j >= 0; --j) {
            t[j] *= dinv[n0 + j];
            for (int k = 0; k < j; ++k) t[k] -= D[tri(j) + k] * t[j];
        }
    }
    for (int j = n0 - 1; j >= 0; --j) {
        double acc = 0.0;
        for (int p = S.Lp[j] + 1; p < S.Lp[j + 1]; ++p) acc += L[p] * y[S.Li[p]];
        y[j] = (y[j] - acc) * dinv[j];
    }
    for (int k = 0; k < n; ++k) x[S.perm[k]] = y[k];
    if (stats) {
        stats[0] = S.nnzL; stats[1] = S.nlev; stats[2] = S.flops; stats[3] = (int64_t)S.as_abr.size() / 4;
        stats[4] = S.T; stats[5] = S.nlev_total;
    }
    return 0;
}

extern "C" int symcheck_solve(int n, int m, const int* Jrp, const int* Jcol, const double* Jv, const int* Prp, const int* Pcol,
                              const double* Pv, const double* d, const double* w, const double* rhs, double* x, int64_t* stats) {
    int64_t st[6];
    int rc = symcheck_solve2(n, m, Jrp, Jcol, Jv, Prp, Pcol, Pv, d, w, rhs, x, stats ? st : nullptr, 0);
    if (stats && rc == 0) for (int k = 0; k < 4; ++k) stats[k] = st[k];
    return rc;
}
