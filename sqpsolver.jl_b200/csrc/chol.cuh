// chol.cuh -- batched sparse Cholesky of the condensed Newton matrix, executed by one team.
//
//      K = P + diag(d) + J' diag(w) J   ->   K = L L'
//
// The symbolic analysis (symbolic.hpp: minimum-degree ordering, level-major column numbering,
// dense tail, and the assembly / factor / solve index programs) is shared by all instances of a
// batch and lives in HBM/L2 as int32 arrays; each instance owns nnz(L) values in global memory
// plus -- CTA teams -- the dense tail of the factor, the inverse diagonal and the solve scratch in
// shared memory.
//
//   sparse levels (columns 0..n0-1): one barrier phase per elimination-tree level; a sub-warp owns
//       a column, forms its diagonal cooperatively and then its sub-diagonal entries lane by lane.
//       Entries are stored in execution order, so the only dependent chain per phase is
//       column pointer -> pair list -> L values.
//   dense tail (last T columns = the chain at the top of the tree, where L is nearly dense):
//       its Schur complement is gathered in ONE parallel phase into a packed T x T lower triangle
//       in shared memory and factorised there by a right-looking dense Cholesky (2 cheap
//       shared-memory barriers per column, no global traffic); the triangular solves on it run in
//       registers of one warp with shuffle broadcasts.
// Deterministic (no atomics); a non-positive pivot is reported, never hidden.
#pragma once
#include "team.cuh"

// lanes (power of two <= 32) that cooperate on one column/row of a level with `count`
// independent items: wide when the level is narrow, 1 when the level has at least as many items
// as the team has threads.
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

// L <- lower triangle of K in the permuted order (all entries, tail included).  Pv may be null
// (no P); dg[j] is added to the diagonal entry of ORIGINAL column j.  Also clears the dense tail.
template <class Team>
__device__ void chol_assemble(Team& T, const CholDev& C, const CholWork& W, const double* __restrict__ Pv,
                              const double* __restrict__ dg, const double* __restrict__ w, const double* __restrict__ Jv) {
    double* __restrict__ L = W.L;
    for (int e = T.tid(); e < C.nnzL; e += T.size()) {
        const int4 hd = C.as_hd[e];
        double v = 0.0;
        if (Pv && hd.x >= 0) v += Pv[hd.x];
        if (hd.y >= 0) v += dg[hd.y];
        for (int t = hd.z; t < hd.w; ++t) {
            const int4 abr = C.as_abr[t];
            v = fma(w[abr.z] * Jv[abr.x], Jv[abr.y], v);
        }
        L[e] = v;
    }
    if (C.T > 0)
        for (int i = T.tid(); i < C.T * (C.T + 1) / 2; i += T.size()) W.D[i] = 0.0;
    T.sync();
}

// ---- dense tail, CTA teams only -------------------------------------------------------------
// packed row-major lower triangle: (r, c), c <= r, at r (r + 1) / 2 + c
__device__ __forceinline__ int tri(int r) { return (r * (r + 1)) >> 1; }

// right-looking dense Cholesky of D (T x T); writes 1/L_jj to dinvT[0..T).  Returns 1.0 if a pivot
// was not positive (every thread sees the same pivots, so the flag is uniform).
__device__ inline double dense_factor(double* __restrict__ D, double* __restrict__ col, double* __restrict__ dinvT, int Tn) {
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, wp = tid >> 5, nw = nth >> 5;
    double bad = 0.0;
    for (int j = 0; j < Tn; ++j) {
        double d = D[tri(j) + j];
        if (!(d > 0.0)) { bad = 1.0; d = 1.0; }
        const double inv = 1.0 / sqrt(d);
        for (int i = j + 1 + tid; i < Tn; i += nth) {
            double c = D[tri(i) + j] * inv;
            D[tri(i) + j] = c;
            col[i] = c;
        }
        if (tid == 0) dinvT[j] = inv;
        __syncthreads();
        for (int i = j + 1 + wp; i < Tn; i += nw) {
            const double ci = col[i];
            double* __restrict__ row = D + tri(i);
            for (int k = j + 1 + lane; k <= i; k += 32) row[k] = fma(-ci, col[k], row[k]);
        }
        __syncthreads();
    }
    return bad;
}

// y <- L_T^{-1} y, then y <- L_T^{-T} y on the tail part of the solve scratch; one warp, the
// right-hand side in registers (lane owns rows lane, lane+32, ...), pivots broadcast by shuffle.
// Tn <= 128.
__device__ inline void dense_solve_warp(const double* __restrict__ D, const double* __restrict__ dinvT, double* yt, int Tn) {
    const int lane = threadIdx.x & 31;
    double t[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) t[s] = (lane + 32 * s < Tn) ? yt[lane + 32 * s] : 0.0;
    // forward, column oriented: y_j = t_j / L_jj ; t_i -= L_ij y_j  (i > j)
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (32 * s >= Tn) break;
        for (int jj = 0; jj < 32; ++jj) {
            const int j = 32 * s + jj;
            if (j >= Tn) break;
            const double yj = __shfl_sync(0xffffffffu, t[s], jj) * dinvT[j];
            if (lane == jj) t[s] = yj;
#pragma unroll
            for (int s2 = 0; s2 < 4; ++s2) {
                if (s2 < s) continue;
                const int i = lane + 32 * s2;
                if (i > j && i < Tn) t[s2] = fma(-D[tri(i) + j], yj, t[s2]);
            }
        }
    }
    // backward, row oriented on L' : x_j = t_j / L_jj ; t_k -= L_jk x_j  (k < j)
#pragma unroll
    for (int s = 3; s >= 0; --s) {
        if (32 * s >= Tn) continue;
        for (int jj = 31; jj >= 0; --jj) {
            const int j = 32 * s + jj;
            if (j >= Tn) continue;
            const double xj = __shfl_sync(0xffffffffu, t[s], jj) * dinvT[j];
            if (lane == jj) t[s] = xj;
            const double* __restrict__ row = D + tri(j);
#pragma unroll
            for (int s2 = 0; s2 < 4; ++s2) {
                if (s2 > s) continue;
                const int k = lane + 32 * s2;
                if (k < j) t[s2] = fma(-row[k], xj, t[s2]);
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s)
        if (lane + 32 * s < Tn) yt[lane + 32 * s] = t[s];
}

// In-place numeric factorisation.  Returns false (uniformly) if a pivot was not positive.
template <class Team>
__device__ bool chol_factor(Team& T, const CholDev& C, const CholWork& W, Prof& pf) {
    double* __restrict__ L = W.L;
    double bad[1] = {0.0};
    for (int l = 0; l < C.nlev; ++l) {
        const int c0 = C.lev_ptr[l], cnt = C.lev_ptr[l + 1] - c0, lg = level_lg(T, cnt), Ln = 1 << lg;
        const int lane = T.tid() & (Ln - 1), sub = T.tid() >> lg, nsub = T.size() >> lg;
        for (int t0 = 0; t0 < cnt; t0 += nsub) {
            const int t = t0 + sub;
            const bool on = t < cnt;
            const int j = c0 + (on ? t : 0);
            const int e0 = C.Lp[j], e1 = on ? C.Lp[j + 1] : e0;
            // diagonal: K_jj - sum_k L_jk^2, pairs spread over the sub-warp
            double acc = 0.0;
            if (on)
                for (int q = C.fp_ptr[e0] + lane, qe = C.fp_ptr[e0 + 1]; q < qe; q += Ln) {
                    double a = L[C.fp_ab[q].x];
                    acc = fma(a, a, acc);
                }
            acc = subwarp_sum(acc, Ln);
            double d = L[e0] - acc;
            if (on && !(d > 0.0)) { bad[0] = 1.0; d = 1.0; }
            const double inv = on ? 1.0 / sqrt(d) : 0.0;
            // sub-diagonal entries of the column, one lane each
            for (int e = e0 + 1 + lane; e < e1; e += Ln) {
                double a2 = 0.0;
                for (int q = C.fp_ptr[e], qe = C.fp_ptr[e + 1]; q < qe; ++q) {
                    const int2 ab = C.fp_ab[q];
                    a2 = fma(L[ab.x], L[ab.y], a2);
                }
                L[e] = (L[e] - a2) * inv;
            }
            __syncwarp();  // every lane has read K_jj before it is overwritten
            if (on && lane == 0) { L[e0] = d * inv; W.dinv[j] = inv; }
        }
        T.sync();
    }
    pf.lap(PS_FACTOR_SPARSE);
    if (C.T > 0) {
        // Schur complement of the tail: S_ij = K_ij - sum_{k < n0} L_ik L_jk, all entries independent
        const int base = C.Lp[C.n0], ne = C.nnzL - base, lg = level_lg(T, ne), Ln = 1 << lg;
        const int lane = T.tid() & (Ln - 1), sub = T.tid() >> lg, nsub = T.size() >> lg;
        for (int t0 = 0; t0 < ne; t0 += nsub) {
            const int t = t0 + sub;
            const bool on = t < ne;
            const int e = base + (on ? t : 0);
            double acc = 0.0;
            if (on)
                for (int q = C.fp_ptr[e] + lane, qe = C.fp_ptr[e + 1]; q < qe; q += Ln) {
                    const int2 ab = C.fp_ab[q];
                    acc = fma(L[ab.x], L[ab.y], acc);
                }
            acc = subwarp_sum(acc, Ln);
            if (on && lane == 0) W.D[C.tpos[t]] = L[e] - acc;
        }
        T.sync();
        pf.lap(PS_SCHUR);
        bad[0] = fmax(bad[0], dense_factor(W.D, W.col, W.dinv + C.n0, C.T));
        pf.lap(PS_FACTOR_DENSE);
    }
    T.template reduce<1, true>(bad);
    return bad[0] == 0.0;
}

// x = K^{-1} b   (b, x in original order; x may alias b)
template <class Team>
__device__ void chol_solve(Team& T, const CholDev& C, const CholWork& W, const double* b, double* x, Prof& pf) {
    const double* __restrict__ L = W.L;
    double* yw = W.yw;
    const double* __restrict__ dinv = W.dinv;
    for (int k = T.tid(); k < C.n; k += T.size()) yw[k] = b[C.perm[k]];
    T.sync();
    for (int l = 0; l < C.nlev; ++l) {  // forward: rows of L
        const int c0 = C.lev_ptr[l], cnt = C.lev_ptr[l + 1] - c0, lg = level_lg(T, cnt), Ln = 1 << lg;
        const int lane = T.tid() & (Ln - 1), sub = T.tid() >> lg, nsub = T.size() >> lg;
        for (int t0 = 0; t0 < cnt; t0 += nsub) {
            const int t = t0 + sub;
            const bool on = t < cnt;
            const int j = c0 + (on ? t : 0);
            double acc = 0.0;
            if (on)
                for (int q = C.Rp[j] + lane, qe = C.Rp[j + 1]; q < qe; q += Ln) {
                    const int2 ic = C.Rci[q];
                    acc = fma(L[ic.x], yw[ic.y], acc);
                }
            acc = subwarp_sum(acc, Ln);
            if (on && lane == 0) yw[j] = (yw[j] - acc) * dinv[j];
        }
        T.sync();
    }
    pf.lap(PS_FWD);
    if (C.T > 0) {
        // tail right-hand side: b_j - sum_{k < n0} L_jk y_k, then the dense solves (one warp)
        const int lg = level_lg(T, C.T), Ln = 1 << lg;
        const int lane = T.tid() & (Ln - 1), sub = T.tid() >> lg, nsub = T.size() >> lg;
        for (int t0 = 0; t0 < C.T; t0 += nsub) {
            const int t = t0 + sub;
            const bool on = t < C.T;
            const int j = C.n0 + (on ? t : 0);
            double acc = 0.0;
            if (on)
                for (int q = C.Rp[j] + lane, qe = C.Rmid[j]; q < qe; q += Ln) {
                    const int2 ic = C.Rci[q];
                    acc = fma(L[ic.x], yw[ic.y], acc);
                }
            acc = subwarp_sum(acc, Ln);
            if (on && lane == 0) yw[j] -= acc;
        }
        T.sync();
        if (threadIdx.x < 32) dense_solve_warp(W.D, dinv + C.n0, yw + C.n0, C.T);
        T.sync();
        pf.lap(PS_TAIL);
    }
    for (int l = C.nlev - 1; l >= 0; --l) {  // backward: columns of L
        const int c0 = C.lev_ptr[l], cnt = C.lev_ptr[l + 1] - c0, lg = level_lg(T, cnt), Ln = 1 << lg;
        const int lane = T.tid() & (Ln - 1), sub = T.tid() >> lg, nsub = T.size() >> lg;
        for (int t0 = 0; t0 < cnt; t0 += nsub) {
            const int t = t0 + sub;
            const bool on = t < cnt;
            const int j = c0 + (on ? t : 0);
            double acc = 0.0;
            if (on)
                for (int p = C.Lp[j] + 1 + lane, pe = C.Lp[j + 1]; p < pe; p += Ln) acc = fma(L[p], yw[C.Li[p]], acc);
            acc = subwarp_sum(acc, Ln);
            if (on && lane == 0) yw[j] = (yw[j] - acc) * dinv[j];
        }
        T.sync();
    }
    for (int k = T.tid(); k < C.n; k += T.size()) x[C.perm[k]] = yw[k];
    T.sync();
    pf.lap(PS_BWD);
}
