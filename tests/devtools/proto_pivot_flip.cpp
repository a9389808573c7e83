// DEVELOPMENT TOOL ONLY: dense LDL' with pivot modification in the device's elimination order (see proto_pivot_flip.py)
#include "../../sqpsolver.jl_b200/csrc/symbolic.hpp"
#include <cmath>
#include <cstdio>
// permutation of the device's symbolic analysis
extern "C" int sym_perm(int n, int m, const int* Jrp, const int* Jcol, const int* Prp, const int* Pcol, int tail_max, int* perm) {
    Symbolic S = symbolic_analyze(n, m, Jrp, Jrp + 1, Jcol, Prp, Pcol, 512, tail_max, true);
    if (!S.ok) return -1;
    for (int k = 0; k < n; ++k) perm[k] = S.perm[k];
    return 0;
}
// dense LDL' (no pivoting) of the symmetric matrix A (row-major n x n, lower part used), in place: L unit lower in the strict
// lower triangle, d on the diagonal.  mode 0: stop at the first pivot <= 0 (returns its index + 1).
// mode 1: a pivot d <= thresh is replaced by max(|d|, floor) ; mode 2: replaced by `floor` ; nmod = number of modified pivots
extern "C" int ldl_mod(double* A, int n, int mode, double thresh, double flo, int* nmod, double* minpiv) {
    *nmod = 0; *minpiv = 1e300;
    for (int j = 0; j < n; ++j) {
        double d = A[(size_t)j * n + j];
        if (d < *minpiv) *minpiv = d;
        if (!(d > thresh)) {
            if (mode == 0) return j + 1;
            double nd = mode == 1 ? std::fmax(std::fabs(d), flo) : flo;
            d = nd; ++*nmod;
        }
        A[(size_t)j * n + j] = d;
        const double inv = 1.0 / d;
        static thread_local std::vector<double> col; col.resize(n);
        for (int i = j + 1; i < n; ++i) { A[(size_t)i * n + j] *= inv; col[i] = A[(size_t)i * n + j]; }   // L_ij
        for (int i = j + 1; i < n; ++i) {
            const double lij = col[i] * d;
            if (lij == 0.0) continue;
            double* Ai = A + (size_t)i * n;
            for (int k = j + 1; k <= i; ++k) Ai[k] -= lij * col[k];
        }
    }
    return 0;
}
extern "C" void ldl_solve(const double* A, int n, double* x) {
    for (int i = 0; i < n; ++i) { double s = x[i]; const double* Ai = A + (size_t)i * n; for (int k = 0; k < i; ++k) s -= Ai[k] * x[k]; x[i] = s; }
    for (int i = 0; i < n; ++i) x[i] /= A[(size_t)i * n + i];
    for (int i = n - 1; i >= 0; --i) { double s = x[i]; for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * x[k]; x[i] = s; }
}
