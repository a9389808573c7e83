// chol.cuh -- batched sparse Cholesky of the condensed Newton matrix, executed by one team.
//
//      K = P + diag(d) + J' diag(w) J   ->   K = L L'
//
// The symbolic analysis (symbolic.hpp: minimum-degree ordering, column structure of L,
// elimination-tree levels and the assembly / factor / solve index programs) is shared by all
// instances of a batch and lives in HBM/L2 as int32 arrays; each instance only owns its nnz(L)
// values (kept in shared memory when they fit).  Every phase is level-scheduled: all columns of
// one elimination-tree level are independent, one team barrier per level and half-phase.
// Deterministic (no atomics); a non-positive pivot is reported, never hidden.
#pragma once
#include "team.cuh"


// L <- lower triangle of K in the permuted order.  Pv may be null (no P); dg[j] is added to the
// diagonal entry of ORIGINAL column j.
template <class Team>
__device__ void chol_assemble(Team& T, const CholDev& C, double* __restrict__ L, const double* __restrict__ Pv,
                              const double* __restrict__ dg, const double* __restrict__ w, const double* __restrict__ Jv) {
    for (int e = T.tid(); e < C.nnzL; e += T.size()) {
        double v = 0.0;
        int h = C.as_h[e], d = C.as_d[e];
        if (Pv && h >= 0) v += Pv[h];
        if (d >= 0) v += dg[d];
        for (int t = C.as_ptr[e]; t < C.as_ptr[e + 1]; ++t) v = fma(w[C.as_r[t]] * Jv[C.as_a[t]], Jv[C.as_b[t]], v);
        L[e] = v;
    }
    T.sync();
}

// lanes (power of two <= 32) that cooperate on one entry/row/column of a level with `count`
// independent items: wide when the level is narrow (top of the elimination tree, long dependent
// gather chains -> spread each chain over a sub-warp so its loads are in flight together),
// 1 when the level has at least as many items as the team has threads.
template <class Team>
__device__ __forceinline__ int level_lg(Team& T, int count) {
    int lg = 0;
    while (lg < 5 && (count << (lg + 1)) <= T.size()) ++lg;
    return lg;
}
__device__ __forceinline__ double subwarp_sum(double v, int L) {
    for (int o = L >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// In-place numeric factorisation.  Returns false (uniformly) if a pivot was not positive.
template <class Team>
__device__ bool chol_factor(Team& T, const CholDev& C, double* __restrict__ L) {
    double bad[1] = {0.0};
    for (int l = 0; l < C.nlev; ++l) {
        const int d0 = C.fd_ptr[l], o0 = C.fo_ptr[l], o1 = C.fd_ptr[l + 1];
        {   // diagonal entries of the level's columns
            const int cnt = o0 - d0, lg = level_lg(T, cnt), Ln = 1 << lg;
            const int lane = T.tid() & (Ln - 1), sub = T.tid() >> lg, nsub = T.size() >> lg;
            for (int t0 = 0; t0 < cnt; t0 += nsub) {
                int t = t0 + sub;
                int e = (t < cnt) ? C.f_ent[d0 + t] : -1;
                double acc = 0.0;
                if (e >= 0)
                    for (int q = C.fp_ptr[e] + lane; q < C.fp_ptr[e + 1]; q += Ln) { double a = L[C.fp_a[q]]; acc = fma(a, a, acc); }
                acc = subwarp_sum(acc, Ln);
                if (e >= 0 && lane == 0) {
                    double v = L[e] - acc;
                    if (!(v > 0.0)) { bad[0] = 1.0; v = 1.0; }
                    L[e] = sqrt(v);
                }
            }
        }
        T.sync();
        {   // off-diagonal entries
            const int cnt = o1 - o0, lg = level_lg(T, cnt), Ln = 1 << lg;
            const int lane = T.tid() & (Ln - 1), sub = T.tid() >> lg, nsub = T.size() >> lg;
            for (int t0 = 0; t0 < cnt; t0 += nsub) {
                int t = t0 + sub;
                int e = (t < cnt) ? C.f_ent[o0 + t] : -1;
                double acc = 0.0;
                if (e >= 0)
                    for (int q = C.fp_ptr[e] + lane; q < C.fp_ptr[e + 1]; q += Ln) acc = fma(L[C.fp_a[q]], L[C.fp_b[q]], acc);
                acc = subwarp_sum(acc, Ln);
                if (e >= 0 && lane == 0) L[e] = (L[e] - acc) / L[C.ent_diag[e]];
            }
        }
        T.sync();
    }
    T.template reduce<1, true>(bad);
    return bad[0] == 0.0;
}

// x = K^{-1} b   (b, x in original order; yw: n-vector of scratch in permuted order; x may alias b)
template <class Team>
__device__ void chol_solve(Team& T, const CholDev& C, const double* __restrict__ L, const double* b, double* x,
                           double* __restrict__ yw) {
    for (int k = T.tid(); k < C.n; k += T.size()) yw[k] = b[C.perm[k]];
    T.sync();
    for (int l = 0; l < C.nlev; ++l) {  // forward: rows of L
        const int c0 = C.lev_ptr[l], cnt = C.lev_ptr[l + 1] - c0, lg = level_lg(T, cnt), Ln = 1 << lg;
        const int lane = T.tid() & (Ln - 1), sub = T.tid() >> lg, nsub = T.size() >> lg;
        for (int t0 = 0; t0 < cnt; t0 += nsub) {
            int t = t0 + sub;
            int j = (t < cnt) ? C.lev_cols[c0 + t] : -1;
            double acc = 0.0;
            if (j >= 0)
                for (int q = C.Rp[j] + lane; q < C.Rp[j + 1]; q += Ln) acc = fma(L[C.Ri[q]], yw[C.Rc[q]], acc);
            acc = subwarp_sum(acc, Ln);
            if (j >= 0 && lane == 0) yw[j] = (yw[j] - acc) / L[C.Lp[j]];
        }
        T.sync();
    }
    for (int l = C.nlev - 1; l >= 0; --l) {  // backward: columns of L
        const int c0 = C.lev_ptr[l], cnt = C.lev_ptr[l + 1] - c0, lg = level_lg(T, cnt), Ln = 1 << lg;
        const int lane = T.tid() & (Ln - 1), sub = T.tid() >> lg, nsub = T.size() >> lg;
        for (int t0 = 0; t0 < cnt; t0 += nsub) {
            int t = t0 + sub;
            int j = (t < cnt) ? C.lev_cols[c0 + t] : -1;
            double acc = 0.0;
            if (j >= 0)
                for (int p = C.Lp[j] + 1 + lane; p < C.Lp[j + 1]; p += Ln) acc = fma(L[p], yw[C.Li[p]], acc);
            acc = subwarp_sum(acc, Ln);
            if (j >= 0 && lane == 0) yw[j] = (yw[j] - acc) / L[C.Lp[j]];
        }
        T.sync();
    }
    for (int k = T.tid(); k < C.n; k += T.size()) x[C.perm[k]] = yw[k];
    T.sync();
}
