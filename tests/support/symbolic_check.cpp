// TEST INFRASTRUCTURE ONLY (not part of libsqpqp.so, never on the product path).
// Executes the index programs produced by csrc/symbolic.hpp on the host, in the order the
// device kernels execute them, so that tests/test_symbolic.py can validate the symbolic
// analysis against a dense Cholesky on a machine without a GPU.
#include <cmath>
#include <cstdint>
#include <vector>
#include "../../sqpsolver.jl_b200/csrc/symbolic.hpp"

extern "C" int symcheck_solve(int n, int m, const int* Jrp, const int* Jcol, const double* Jv, const int* Prp, const int* Pcol,
                              const double* Pv, const double* d, const double* w, const double* rhs, double* x,
                              int64_t* stats /* nnzL, nlev, flops, as_len */) {
    Symbolic S = symbolic_analyze(n, m, Jrp, Jrp + 1, Jcol, Prp, Pcol);
    if (!S.ok) return -1;
    std::vector<double> L(S.nnzL);
    for (int e = 0; e < S.nnzL; ++e) {
        double v = 0.0;
        if (S.as_h[e] >= 0) v += Pv[S.as_h[e]];
        if (S.as_d[e] >= 0) v += d[S.as_d[e]];
        for (int t = S.as_ptr[e]; t < S.as_ptr[e + 1]; ++t) v += w[S.as_r[t]] * Jv[S.as_a[t]] * Jv[S.as_b[t]];
        L[e] = v;
    }
    for (int l = 0; l < S.nlev; ++l) {
        for (int t = S.fd_ptr[l]; t < S.fo_ptr[l]; ++t) {
            int e = S.f_ent[t];
            double v = L[e];
            for (int q = S.fp_ptr[e]; q < S.fp_ptr[e + 1]; ++q) v -= L[S.fp_a[q]] * L[S.fp_b[q]];
            if (!(v > 0.0)) return -2;
            L[e] = std::sqrt(v);
        }
        for (int t = S.fo_ptr[l]; t < S.fd_ptr[l + 1]; ++t) {
            int e = S.f_ent[t];
            double v = L[e];
            for (int q = S.fp_ptr[e]; q < S.fp_ptr[e + 1]; ++q) v -= L[S.fp_a[q]] * L[S.fp_b[q]];
            L[e] = v / L[S.ent_diag[e]];
        }
    }
    std::vector<double> y(n);
    for (int k = 0; k < n; ++k) y[k] = rhs[S.perm[k]];
    for (int l = 0; l < S.nlev; ++l)
        for (int t = S.lev_ptr[l]; t < S.lev_ptr[l + 1]; ++t) {
            int j = S.lev_cols[t];
            double acc = y[j];
            for (int q = S.Rp[j]; q < S.Rp[j + 1]; ++q) acc -= L[S.Ri[q]] * y[S.Rc[q]];
            y[j] = acc / L[S.Lp[j]];
        }
    for (int l = S.nlev - 1; l >= 0; --l)
        for (int t = S.lev_ptr[l]; t < S.lev_ptr[l + 1]; ++t) {
            int j = S.lev_cols[t];
            double acc = y[j];
            for (int p = S.Lp[j] + 1; p < S.Lp[j + 1]; ++p) acc -= L[p] * y[S.Li[p]];
            y[j] = acc / L[S.Lp[j]];
        }
    for (int k = 0; k < n; ++k) x[S.perm[k]] = y[k];
    if (stats) { stats[0] = S.nnzL; stats[1] = S.nlev; stats[2] = S.flops; stats[3] = (int64_t)S.as_a.size(); }
    return 0;
}
