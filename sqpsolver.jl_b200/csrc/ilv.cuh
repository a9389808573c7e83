// ilv.cuh -- interior-point QP solve for the BATCH, G instances interleaved per CTA (values
// batch-innermost).
//
// The CTA team of ipm.cuh / chol.cuh gives one instance to one CTA: every index of the shared symbolic programs is
// loaded once per instance and every gathered factor value uses 8 bytes of a 32-byte sector.  Here a CTA owns a GROUP
// of G instances (G = 2, 4 or 8) that share the sparsity pattern, and every per-instance array of the solve is stored
// group-interleaved,
//
//        a[(group * len + i) * G + g]            g = instance within the group (fastest), i = element
//
// A thread is (task lane tl, instance g): g = threadIdx.x % G, tl = threadIdx.x / G.  The G threads of a task lane
// execute the SAME entry of the index programs on G different instances, so
//   * an index / descriptor load is shared by G lanes (one transaction),
//   * a value load or gather is G consecutive doubles: full 32-byte sectors from G = 4 up,
//   * a CTA barrier and the dependent-load chain behind it are paid once per G instances,
//   * lanes of a multi-lane task sit at stride G in the warp (shuffle offsets o * G).
// The algorithm is the one of ipm.cuh (same passes, same barrier rule, same inertia correction, same termination
// tests); what differs is control flow: the G instances of a group iterate in lock step, so every per-instance
// decision (converged, infeasible, factorisation failed) is a PREDICATE (`live`, `need`) on the stores instead of a
// branch around barriers, and a loop runs while any instance of the group needs it (__syncthreads_or).
// An instance the interior point cannot finish is flagged in P.fb_flag exactly like MODE 1 of k_solve_cta does, and
// the masked ADMM launch that follows picks it up (admm.cuh).
//
// Replaces, for a batch: the external solve behind JuMP.optimize! in sub_optimize! / sub_optimize_FR! /
// sub_optimize_lp (subproblem_JuMP.jl:178, 388, 209) with set_trust_region! (:432-448), modify_constraints! (:465-512)
// and collect_solution! (:514-563).
#pragma once
#include "ipm.cuh"

struct IlvDev {
    CholDev C;                       // symbolic program with the slot lists built for 32 / G lanes per task
    const double *Jvi, *Tvi, *Hvi;   // unscaled matrix values, group-interleaved (written by k_scatter)
    int off_D, off_yw, off_dinv;     // offsets (doubles) into dynamic shared memory; -1: the array stays in global memory
    int ngroups;
};

#define ILV_KMAX 8

template <int G>
struct IlvTeam {
    static constexpr int LG = (G == 1) ? 0 : (G == 2) ? 1 : (G == 4) ? 2 : 3;
    static constexpr int W = 32 / G;        // task lanes per warp
    static constexpr int MAXLG = 5 - LG;    // lanes (log2) one task can use
    double* sh;                             // [2][ILV_KMAX][warps of the CTA][G]
    int flip;
    int g, tl, TL;
    __device__ IlvTeam(double* s) : sh(s), flip(0) {
        g = threadIdx.x & (G - 1);
        tl = threadIdx.x >> LG;
        TL = blockDim.x >> LG;
    }
    // per-instance reduction over the task lanes: result in v[] for every thread of that instance.  Deterministic.
    template <int K, bool IS_MAX>
    __device__ void reduce(double (&v)[K]) {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        double* buf = sh + flip * (ILV_KMAX * nw * G);
        flip ^= 1;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double r = v[k];
#pragma unroll
            for (int o = 16; o >= G; o >>= 1) {
                const double t = __shfl_xor_sync(0xffffffffu, r, o);
                r = IS_MAX ? fmax(r, t) : r + t;
            }
            if (lane < G) buf[(k * nw + w) * G + lane] = r;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double r = IS_MAX ? -INFINITY : 0.0;
            for (int w0 = 0; w0 < nw; w0 += W) {  // lane reads warp (w0 + lane / G), instance lane % G: consecutive addresses
                const int ww = w0 + (lane >> LG);
                const double t = (ww < nw) ? buf[(k * nw + ww) * G + g] : (IS_MAX ? -INFINITY : 0.0);
                r = IS_MAX ? fmax(r, t) : r + t;
            }
#pragma unroll
            for (int o = 16; o >= G; o >>= 1) {
                const double t = __shfl_xor_sync(0xffffffffu, r, o);
                r = IS_MAX ? fmax(r, t) : r + t;
            }
            v[k] = r;
        }
        // double-buffered like CtaTeam::reduce: no trailing barrier needed
    }
};

// sum over the 2^lg task lanes of a lane group (lanes at stride G in the warp)
template <int G>
__device__ __forceinline__ double ilv_group_sum(double v, int Ln) {
    for (int o = Ln >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o * G);
    return v;
}

// CSR row products on interleaved values: 2^lg task lanes per row.  f(row, dot) by lane 0 of the group.
template <int G, class F>
__device__ __forceinline__ void ilv_rows(const IlvTeam<G>& T, int nrows, int lg, const int* __restrict__ rb, const int* __restrict__ re,
                                         const int* __restrict__ col, const double* __restrict__ val, const double* __restrict__ x, F f) {
    const int L = 1 << lg, lane = T.tl & (L - 1), sub = T.tl >> lg, nsub = T.TL >> lg;
    for (int r0 = 0; r0 < nrows; r0 += nsub) {
        const int r = r0 + sub;
        double acc = 0.0;
        if (r < nrows) {
            const int e = re[r];
            for (int k = rb[r] + lane; k < e; k += L) acc = fma(val[(size_t)k * G], x[(size_t)col[k] * G], acc);
        }
        acc = ilv_group_sum<G>(acc, L);
        if (r < nrows && lane == 0) f(r, acc);
    }
}
template <int G, class F>
__device__ __forceinline__ void ilv_rows2(const IlvTeam<G>& T, int nrows, int lg, const int* __restrict__ rb1, const int* __restrict__ re1,
                                          const int* __restrict__ col1, const double* __restrict__ val1, const double* __restrict__ x1,
                                          bool use1, const int* __restrict__ rb2, const int* __restrict__ re2,
                                          const int* __restrict__ col2, const double* __restrict__ val2, const double* __restrict__ x2, F f) {
    const int L = 1 << lg, lane = T.tl & (L - 1), sub = T.tl >> lg, nsub = T.TL >> lg;
    for (int r0 = 0; r0 < nrows; r0 += nsub) {
        const int r = r0 + sub;
        double a1 = 0.0, a2 = 0.0;
        if (r < nrows) {
            if (use1) {
                const int e = re1[r];
                for (int k = rb1[r] + lane; k < e; k += L) a1 = fma(val1[(size_t)k * G], x1[(size_t)col1[k] * G], a1);
            }
            const int e2 = re2[r];
            for (int k = rb2[r] + lane; k < e2; k += L) a2 = fma(val2[(size_t)k * G], x2[(size_t)col2[k] * G], a2);
        }
        for (int o = L >> 1; o > 0; o >>= 1) {
            a1 += __shfl_xor_sync(0xffffffffu, a1, o * G);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o * G);
        }
        if (r < nrows && lane == 0) f(r, a1, a2);
    }
}

// sum_{q = q0, q0 + step, ... < qe} A[ab[q].x * G] * B[ab[q].y * G], four independent gather chains in flight
template <int G>
__device__ __forceinline__ double ilv_gather_dot(const int2* __restrict__ ab, int q, const int qe, const int step, const double* A,
                                                 const double* B) {
    double acc0 = 0.0, acc1 = 0.0;
    for (; q + 3 * step < qe; q += 4 * step) {
        const int2 p0 = ab[q], p1 = ab[q + step], p2 = ab[q + 2 * step], p3 = ab[q + 3 * step];
        const double a0 = A[(size_t)p0.x * G], b0 = B[(size_t)p0.y * G], a1 = A[(size_t)p1.x * G], b1 = B[(size_t)p1.y * G];
        const double a2 = A[(size_t)p2.x * G], b2 = B[(size_t)p2.y * G], a3 = A[(size_t)p3.x * G], b3 = B[(size_t)p3.y * G];
        acc0 = fma(a0, b0, acc0); acc1 = fma(a1, b1, acc1);
        acc0 = fma(a2, b2, acc0); acc1 = fma(a3, b3, acc1);
    }
    for (; q < qe; q += step) {
        const int2 p0 = ab[q];
        acc0 = fma(A[(size_t)p0.x * G], B[(size_t)p0.y * G], acc0);
    }
    return acc0 + acc1;
}

__device__ __forceinline__ int ilv_level_lg(int TL, int count, int maxlg) {
    int lg = 0;
    while (lg < maxlg && (count << (lg + 1)) <= TL) ++lg;
    return lg;
}

// per-thread view of the factorisation state of (group, g)
struct IlvChol {
    double* L;      // [nnzL][G] + g
    double* D;      // [Tp(Tp+1)/2][G] + g   (shared memory)
    double* dinv;   // [n][G] + g
    double* yw;     // [n][G] + g
    double* wJ;     // [nslotJ][G] + g
};

// K = P + diag(dg + shift) + J' diag(w) J  ->  L (sourced entries), tail cleared.  Stores predicated by `on`.
template <int G>
__device__ void ilv_assemble(const IlvTeam<G>& T, const CholDev& C, const IlvChol& W, const double* __restrict__ Pv,
                             const double* __restrict__ dg, const double shift, const double* __restrict__ w,
                             const double* __restrict__ Jv, const bool on) {
    for (int a = T.tl; a < C.nslotJ; a += T.TL) {
        const int r = C.jrow[a];
        if (on) W.wJ[(size_t)a * G] = (r >= 0) ? w[(size_t)r * G] * Jv[(size_t)a * G] : 0.0;
    }
    if (C.T > 0) {
        const int Tp = (C.T + 3) & ~3;
        if (on) for (int i = T.tl; i < Tp * (Tp + 1) / 2; i += T.TL) W.D[(size_t)i * G] = 0.0;
        __syncthreads();
        if (on) for (int i = C.T + T.tl; i < Tp; i += T.TL) W.D[(size_t)(i * (i + 1) / 2 + i) * G] = 1.0;
    }
    __syncthreads();
    for (int r0 = 0; r0 < C.n_aslot; r0 += T.TL) {
        const int s = r0 + T.tl;
        const bool act = s < C.n_aslot;
        const int4 sl = C.aslot[act ? s : 0];
        const int Ln = 1 << ((sl.x >> 26) & 7);
        const bool ld = act && ((sl.x >> 29) & 1);
        double v = 0.0;
        if (ld) {
            const int d = C.aslot_d[s];
            if (Pv && sl.w >= 0) v = Pv[(size_t)sl.w * G];
            if (d >= 0) v += dg[(size_t)d * G] + shift;
        }
        double acc = act ? ilv_gather_dot<G>(C.as_ab, sl.y, sl.z, Ln, W.wJ, Jv) : 0.0;
#pragma unroll
        for (int o = IlvTeam<G>::W >> 1; o > 0; o >>= 1) {  // lane groups of mixed (power of two, aligned) sizes share the butterfly
            const double t = __shfl_xor_sync(0xffffffffu, acc, o * G);
            if (o < Ln) acc += t;
        }
        if (ld && on) W.L[(size_t)(sl.x & 0x3ffffff) * G] = v + acc;
    }
    __syncthreads();
}

__device__ __forceinline__ int ilv_tri(int r) { return (r * (r + 1)) >> 1; }

// Right-looking dense Cholesky of the packed tails of the G instances, panels of 4 columns (the algorithm of
// dense_factor in chol.cuh): every thread factorises the 4 x 4 diagonal block of ITS instance in registers, task lane
// r forward-substitutes row r of the panel; the rank-4 update gives a row to a warp and W columns to its task lanes.
template <int G>
__device__ double ilv_dense_factor(const IlvTeam<G>& T, double* __restrict__ D, double* __restrict__ dinvT, int Tn, const bool on) {
    constexpr int W = IlvTeam<G>::W;
    const int wp = threadIdx.x >> 5, nw = blockDim.x >> 5, lt = T.tl & (W - 1);
    const int Tp = (Tn + 3) & ~3;
    double bad = 0.0;
#define DD(idx) D[(size_t)(idx) * G]
    for (int j0 = 0; j0 < Tp; j0 += 4) {
        const int t0 = ilv_tri(j0) + j0, t1 = ilv_tri(j0 + 1) + j0, t2 = ilv_tri(j0 + 2) + j0, t3 = ilv_tri(j0 + 3) + j0;
        double a00 = DD(t0), a10 = DD(t1), a11 = DD(t1 + 1), a20 = DD(t2), a21 = DD(t2 + 1), a22 = DD(t2 + 2);
        double a30 = DD(t3), a31 = DD(t3 + 1), a32 = DD(t3 + 2), a33 = DD(t3 + 3);
        if (!(a00 > 0.0)) { bad = 1.0; a00 = 1.0; }
        const double i0 = rsqrt(a00);
        const double l10 = a10 * i0, l20 = a20 * i0, l30 = a30 * i0;
        a11 = fma(-l10, l10, a11);
        if (!(a11 > 0.0)) { bad = 1.0; a11 = 1.0; }
        const double i1 = rsqrt(a11);
        const double l21 = fma(-l20, l10, a21) * i1, l31 = fma(-l30, l10, a31) * i1;
        a22 = fma(-l21, l21, fma(-l20, l20, a22));
        if (!(a22 > 0.0)) { bad = 1.0; a22 = 1.0; }
        const double i2 = rsqrt(a22);
        const double l32 = fma(-l31, l21, fma(-l30, l20, a32)) * i2;
        a33 = fma(-l32, l32, fma(-l31, l31, fma(-l30, l30, a33)));
        if (!(a33 > 0.0)) { bad = 1.0; a33 = 1.0; }
        const double i3 = rsqrt(a33);
        for (int i = j0 + 4 + T.tl; i < Tp; i += T.TL) {
            const int ri = ilv_tri(i) + j0;
            const double x0 = DD(ri) * i0;
            const double x1 = fma(-x0, l10, DD(ri + 1)) * i1;
            const double x2 = fma(-x1, l21, fma(-x0, l20, DD(ri + 2))) * i2;
            const double x3 = fma(-x2, l32, fma(-x1, l31, fma(-x0, l30, DD(ri + 3)))) * i3;
            if (on) { DD(ri) = x0; DD(ri + 1) = x1; DD(ri + 2) = x2; DD(ri + 3) = x3; }
        }
        __syncthreads();  // every thread has read the diagonal block; the panel rows are written
        if (T.tl == 0 && on) {
            DD(t1) = l10; DD(t2) = l20; DD(t2 + 1) = l21; DD(t3) = l30; DD(t3 + 1) = l31; DD(t3 + 2) = l32;
            if (j0 < Tn) dinvT[(size_t)j0 * G] = i0;
            if (j0 + 1 < Tn) dinvT[(size_t)(j0 + 1) * G] = i1;
            if (j0 + 2 < Tn) dinvT[(size_t)(j0 + 2) * G] = i2;
            if (j0 + 3 < Tn) dinvT[(size_t)(j0 + 3) * G] = i3;
        }
        for (int i = j0 + 4 + wp; i < Tp; i += nw) {  // rank-4 update: a warp per row, W columns per trip
            const int ri = ilv_tri(i);
            const double xi0 = DD(ri + j0), xi1 = DD(ri + j0 + 1), xi2 = DD(ri + j0 + 2), xi3 = DD(ri + j0 + 3);
            for (int k = j0 + 4 + lt; k <= i; k += W) {
                const int rk = ilv_tri(k) + j0;
                const double s = fma(xi3, DD(rk + 3), fma(xi2, DD(rk + 2), fma(xi1, DD(rk + 1), xi0 * DD(rk))));
                if (on) DD(ri + k) -= s;
            }
        }
        __syncthreads();
    }
#undef DD
    return bad;
}

// y <- L_T^{-1} y, y <- L_T^{-T} y on the tail of instance gi; ONE WARP per instance (registers + shuffles, the
// algorithm of dense_solve_warp in chol.cuh on strided storage).  Tn <= 128.
template <int G>
__device__ inline void ilv_dense_solve_warp(const double* __restrict__ D, const double* __restrict__ dinvT, double* yt, int Tn, int gi) {
    const int lane = threadIdx.x & 31;
#define DD(idx) D[(size_t)(idx) * G + gi]
    double t[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) t[s] = (lane + 32 * s < Tn) ? yt[(size_t)(lane + 32 * s) * G + gi] : 0.0;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        if (32 * s < Tn) {
            const int jend = (Tn - 32 * s) < 32 ? (Tn - 32 * s) : 32;
            for (int jj = 0; jj < jend; ++jj) {
                const int j = 32 * s + jj;
                double dl[4];
#pragma unroll
                for (int s2 = 0; s2 < 4; ++s2) {
                    const int i = lane + 32 * s2;
                    dl[s2] = (s2 >= s && i > j && i < Tn) ? DD(ilv_tri(i) + j) : 0.0;
                }
                const double yj = __shfl_sync(0xffffffffu, t[s], jj) * dinvT[(size_t)j * G + gi];
                if (lane == jj) t[s] = yj;
#pragma unroll
                for (int s2 = 0; s2 < 4; ++s2)
                    if (s2 >= s) t[s2] = fma(-dl[s2], yj, t[s2]);
            }
        }
    }
#pragma unroll
    for (int s = 3; s >= 0; --s) {
        if (32 * s < Tn) {
            const int jend = (Tn - 32 * s) < 32 ? (Tn - 32 * s) : 32;
            for (int jj = jend - 1; jj >= 0; --jj) {
                const int j = 32 * s + jj;
                const int rj = ilv_tri(j);
                double dl[4];
#pragma unroll
                for (int s2 = 0; s2 < 4; ++s2) {
                    const int k = lane + 32 * s2;
                    dl[s2] = (s2 <= s && k < j) ? DD(rj + k) : 0.0;
                }
                const double xj = __shfl_sync(0xffffffffu, t[s], jj) * dinvT[(size_t)j * G + gi];
                if (lane == jj) t[s] = xj;
#pragma unroll
                for (int s2 = 0; s2 < 4; ++s2)
                    if (s2 <= s) t[s2] = fma(-dl[s2], xj, t[s2]);
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s)
        if (lane + 32 * s < Tn) yt[(size_t)(lane + 32 * s) * G + gi] = t[s];
#undef DD
}

// numeric factorisation of the G instances (stores predicated by `on`); returns true where no pivot was non-positive
template <int G>
__device__ bool ilv_factor(IlvTeam<G>& T, const CholDev& C, const IlvChol& W, const bool on) {
    double* L = W.L;
    double bad[1] = {0.0};
    int4 ph = C.fphase[0];
    for (int p = 0; p < C.nphase; ++p) {
        const int4 nxt = C.fphase[p + 1];  // (padded by one entry)
        const int s0 = ph.x, ns = ph.y - ph.x, kind = ph.w;
        const bool wide = (ph.z >> 24) > 0;  // some task of the phase has more than one lane
        for (int r0 = 0; r0 < ns; r0 += T.TL) {
            const int s = r0 + T.tl;
            const bool act = s < ns;
            const int4 sl = C.ftask[s0 + (act ? s : 0)];
            const int e = sl.x & 0x3ffffff;
            const int Ln = 1 << ((sl.x >> 26) & 7);
            const bool ld = act && ((sl.x >> 29) & 1);
            const double kv = (ld && (sl.x >> 30)) ? L[(size_t)e * G] : 0.0;
            const double dv = (ld && kind == 1) ? W.dinv[(size_t)sl.w * G] : 0.0;
            double acc = act ? ilv_gather_dot<G>(C.fp_ab, sl.y, sl.z, Ln, L, L) : 0.0;
            if (wide) {
#pragma unroll
                for (int o = IlvTeam<G>::W >> 1; o > 0; o >>= 1) {
                    const double t = __shfl_xor_sync(0xffffffffu, acc, o * G);
                    if (o < Ln) acc += t;
                }
            }
            if (ld) {
                double v = kv - acc;
                if (kind == 0) {
                    if (!(v > 0.0)) { bad[0] = 1.0; v = 1.0; }
                    const double inv = rsqrt(v);
                    if (on) { L[(size_t)e * G] = v * inv; W.dinv[(size_t)sl.w * G] = inv; }
                } else if (kind == 1) {
                    if (on) L[(size_t)e * G] = v * dv;
                } else {
                    if (on) W.D[(size_t)sl.w * G] = v;
                }
            }
        }
        __syncthreads();
        ph = nxt;
    }
    if (C.T > 0) bad[0] = fmax(bad[0], ilv_dense_factor<G>(T, W.D, W.dinv + (size_t)C.n0 * G, C.T, on));
    T.template reduce<1, true>(bad);
    return bad[0] == 0.0;
}

// x = K^{-1} b for the G instances (b, x: interleaved N-vectors in original order); stores predicated by `on`
template <int G>
__device__ void ilv_solve(const IlvTeam<G>& T, const CholDev& C, const IlvChol& W, const double* b, double* x, const bool on) {
    constexpr int MAXLG = IlvTeam<G>::MAXLG;
    const double* L = W.L;
    double* yw = W.yw;
    const double* dinv = W.dinv;
    for (int k = T.tl; k < C.n; k += T.TL) yw[(size_t)k * G] = b[(size_t)C.perm[k] * G];
    __syncthreads();
    for (int l = 0; l < C.nlev; ++l) {  // forward: rows of L
        const int c0 = C.lev_ptr[l], cnt = C.lev_ptr[l + 1] - c0, lg = ilv_level_lg(T.TL, cnt, MAXLG), Ln = 1 << lg;
        const int lane = T.tl & (Ln - 1), sub = T.tl >> lg, nsub = T.TL >> lg;
        for (int t0 = 0; t0 < cnt; t0 += nsub) {
            const int t = t0 + sub;
            const bool act = t < cnt;
            const int j = c0 + (act ? t : 0);
            const double yj = yw[(size_t)j * G], dj = dinv[(size_t)j * G];
            double acc = act ? ilv_gather_dot<G>(C.Rci, C.Rp[j] + lane, C.Rp[j + 1], Ln, L, yw) : 0.0;
            acc = ilv_group_sum<G>(acc, Ln);
            if (act && lane == 0) yw[(size_t)j * G] = (yj - acc) * dj;
        }
        __syncthreads();
    }
    if (C.T > 0) {
        const int lg = ilv_level_lg(T.TL, C.T, MAXLG), Ln = 1 << lg;
        const int lane = T.tl & (Ln - 1), sub = T.tl >> lg, nsub = T.TL >> lg;
        for (int t0 = 0; t0 < C.T; t0 += nsub) {
            const int t = t0 + sub;
            const bool act = t < C.T;
            const int j = C.n0 + (act ? t : 0);
            double acc = act ? ilv_gather_dot<G>(C.Rci, C.Rp[j] + lane, C.Rmid[j], Ln, L, yw) : 0.0;
            acc = ilv_group_sum<G>(acc, Ln);
            if (act && lane == 0) yw[(size_t)j * G] -= acc;
        }
        __syncthreads();
        {   // one warp per instance; base pointers WITHOUT this thread's g
            const int wp = threadIdx.x >> 5;
            if (wp < G)
                ilv_dense_solve_warp<G>(W.D - T.g, dinv - T.g + (size_t)C.n0 * G, yw - T.g + (size_t)C.n0 * G, C.T, wp);
        }
        __syncthreads();
    }
    for (int l = C.nlev - 1; l >= 0; --l) {  // backward: columns of L
        const int c0 = C.lev_ptr[l], cnt = C.lev_ptr[l + 1] - c0, lg = ilv_level_lg(T.TL, cnt, MAXLG), Ln = 1 << lg;
        const int lane = T.tl & (Ln - 1), sub = T.tl >> lg, nsub = T.TL >> lg;
        for (int t0 = 0; t0 < cnt; t0 += nsub) {
            const int t = t0 + sub;
            const bool act = t < cnt;
            const int j = c0 + (act ? t : 0);
            const double yj = yw[(size_t)j * G], dj = dinv[(size_t)j * G];
            double acc0 = 0.0, acc1 = 0.0;
            if (act) {
                int p = C.Lp[j] + 1 + lane;
                const int pe = C.Lp[j + 1];
                for (; p + Ln < pe; p += 2 * Ln) {
                    const int i0 = C.Li[p], i1 = C.Li[p + Ln];
                    acc0 = fma(L[(size_t)p * G], yw[(size_t)i0 * G], acc0);
                    acc1 = fma(L[(size_t)(p + Ln) * G], yw[(size_t)i1 * G], acc1);
                }
                if (p < pe) acc0 = fma(L[(size_t)p * G], yw[(size_t)C.Li[p] * G], acc0);
            }
            const double acc = ilv_group_sum<G>(acc0 + acc1, Ln);
            if (act && lane == 0) yw[(size_t)j * G] = (yj - acc) * dj;
        }
        __syncthreads();
    }
    if (on) for (int k = T.tl; k < C.n; k += T.TL) x[(size_t)C.perm[k] * G] = yw[(size_t)k * G];
    __syncthreads();
}

// ---- the solve of one group --------------------------------------------------------------------------------------
template <int G>
__device__ void ilv_solve_group(IlvTeam<G>& T, const Prob& P, const sqpqp_options& o, const IlvDev& X, const int grp, const int phase,
                                double* dsm) {
    constexpr int MAXLG = IlvTeam<G>::MAXLG;
    const CholDev& C = X.C;
    const int g = T.g, TL = T.TL, tl = T.tl;
    const int n = P.n, m = P.m;
    const bool fr = phase == SQPQP_PHASE_FR;
    const int N = fr ? P.Ne : n, M = m;
    const bool useH = (phase == SQPQP_PHASE_QP || phase == SQPQP_PHASE_SOC) && P.has_hess;
    const double pconst = (phase == SQPQP_PHASE_LP) ? 2.0 : 0.0;
    const int inst_raw = grp * G + g;
    const bool valid = inst_raw < P.batch && (!P.active || P.active[inst_raw < P.batch ? inst_raw : 0]);
    const int inst = inst_raw < P.batch ? inst_raw : P.batch - 1;  // padding lanes read instance batch-1, store nowhere
    // group-interleaved views (+ g)
    const size_t gN = (size_t)grp * G * P.Ne + g, gM = (size_t)grp * G * (m > 0 ? m : 1) + g;
#define NV(k) (P.nv[k] + gN)
#define MV(k) (P.mv[k] + gM)
#define AT(i) [(size_t)(i) * G]
    const int* Jrb = P.J_rb; const int* Jre = fr ? P.J_re_e : P.J_re_n; const int* Jcol = P.J_col;
    const int* Trb = P.T_rb; const int* Tre = P.T_rb + 1; const int* Tcol = P.T_col;
    const int* Hrb = P.H_rb; const int* Hre = P.H_rb + 1; const int* Hcol = P.H_col;
    const int lgJ = min(fr ? P.lgJe : P.lgJn, MAXLG), lgT = min(P.lgT > P.lgH ? P.lgT : P.lgH, MAXLG);
    const double* Jv = X.Jvi + (size_t)grp * G * P.nnzJ + g;
    const double* Tv = X.Tvi + (size_t)grp * G * P.nnzT + g;
    const double* Hv = X.Hvi + (size_t)grp * G * P.nnzH + g;
    double* Jsv = P.Jsv + (size_t)grp * G * P.nnzJ + g;
    double* Tsv = P.Tsv + (size_t)grp * G * P.nnzT + g;
    double* Hsv = P.Hsv + (size_t)grp * G * P.nnzH + g;
    const double* df = P.df + (size_t)inst * n;
    const double* Ecur = (phase == SQPQP_PHASE_SOC && P.Eov) ? P.Eov + (size_t)inst * m : P.E + (size_t)inst * m;
    const double* gL = P.gL + (size_t)inst * P.gstride;
    const double* gU = P.gU + (size_t)inst * P.gstride;
    const double* xL = P.xL + (size_t)inst * P.xstride;
    const double* xU = P.xU + (size_t)inst * P.xstride;
    const double* xk = P.xk + (size_t)inst * n;
    const double delta_tr = P.delta[inst];

    IlvChol W;
    W.L = (fr ? P.Lval_fr : P.Lval) + (size_t)grp * G * C.nnzL + g;
    W.wJ = P.wJ + (size_t)grp * G * P.nnzJ + g;
    W.yw = (X.off_yw >= 0) ? dsm + X.off_yw + g : (fr ? P.yw_fr + (size_t)grp * G * P.Ne : P.yw + (size_t)grp * G * P.n) + g;
    W.dinv = (X.off_dinv >= 0) ? dsm + X.off_dinv + g : (fr ? P.dinv_fr + (size_t)grp * G * P.Ne : P.dinv + (size_t)grp * G * P.n) + g;
    W.D = (X.off_D >= 0) ? dsm + X.off_D + g : nullptr;

    // ================= stage A: QP data (set_trust_region!, modify_constraints!), Ruiz equilibration =================
    double c = 1.0;
    {
        double *q = NV(N_Q), *xl = NV(N_XL), *xu = NV(N_XU), *D = NV(N_D), *hd = NV(N_HD), *tmpN = NV(N_TMP);
        double *rl = MV(M_RL), *ru = MV(M_RU), *Es = MV(M_ES), *tmpM = MV(M_TMP);
        for (int j = tl; j < N; j += TL) {
            double lo, hi, qq;
            if (j < n) {
                if (phase == SQPQP_PHASE_LP) {
                    lo = xL[j]; hi = xU[j]; qq = -2.0 * xk[j];
                } else {
                    const double vl = xL[j] - xk[j], vu = xU[j] - xk[j];
                    lo = fmax(-delta_tr, vl); hi = fmin(delta_tr, vu);
                    if (lo > hi) {  // x_k outside its bounds (subproblem_JuMP.jl:441-444)
                        lo = fmax(-delta_tr, fmin(0.0, vl));
                        hi = fmin(delta_tr, fmax(0.0, vu));
                    }
                    qq = fr ? 0.0 : df[j];
                }
            } else {  // FR slack column: free >= 0 unless its row is already satisfied (:365-380)
                const int i = P.slack_row[j - n];
                const bool feas = (Ecur[i] >= gL[i]) && (Ecur[i] <= gU[i]);
                lo = 0.0; hi = feas ? 0.0 : INFINITY; qq = 1.0;
            }
            xl AT(j) = lo; xu AT(j) = hi; q AT(j) = qq; D AT(j) = 1.0;
        }
        for (int i = tl; i < M; i += TL) {
            double lo, hi;
            if (phase == SQPQP_PHASE_LP) {
                if (i < P.mlin) { lo = gL[i]; hi = gU[i]; } else { lo = -INFINITY; hi = INFINITY; }
            } else {
                lo = gL[i] - Ecur[i]; hi = gU[i] - Ecur[i];
            }
            rl AT(i) = lo; ru AT(i) = hi; Es AT(i) = 1.0;
        }
        __syncthreads();
        for (int it = 0; it < o.ruiz_iters; ++it) {
            for (int j = tl; j < N; j += TL) {
                double a = 0.0;
                if (useH)
                    for (int k = Hrb[j]; k < Hre[j]; ++k) a = fmax(a, fabs(Hv AT(k)) * D AT(Hcol[k]));
                a = c * a;
                if (pconst != 0.0) a = fmax(a, c * pconst * D AT(j));
                double b = 0.0;
                for (int k = Trb[j]; k < Tre[j]; ++k) b = fmax(b, fabs(Tv AT(k)) * Es AT(Tcol[k]));
                const double cn = D AT(j) * fmax(a, b);
                tmpN AT(j) = 1.0 / sqrt((cn > 1e-4) ? fmin(cn, 1e4) : 1.0);
            }
            for (int i = tl; i < M; i += TL) {
                double a = 0.0;
                for (int k = Jrb[i]; k < Jre[i]; ++k) a = fmax(a, fabs(Jv AT(k)) * D AT(Jcol[k]));
                const double rn = Es AT(i) * a;
                tmpM AT(i) = 1.0 / sqrt((rn > 1e-4) ? fmin(rn, 1e4) : 1.0);
            }
            __syncthreads();
            for (int j = tl; j < N; j += TL) D AT(j) *= tmpN AT(j);
            for (int i = tl; i < M; i += TL) Es AT(i) *= tmpM AT(i);
            __syncthreads();
            double s[1] = {0.0}, mx[1] = {0.0};
            for (int j = tl; j < N; j += TL) {
                double a = 0.0;
                if (useH)
                    for (int k = Hrb[j]; k < Hre[j]; ++k) a = fmax(a, fabs(Hv AT(k)) * D AT(Hcol[k]));
                if (pconst != 0.0) a = fmax(a, pconst * D AT(j));
                s[0] += c * D AT(j) * a;
                mx[0] = fmax(mx[0], fabs(c * D AT(j) * q AT(j)));
            }
            T.template reduce<1, false>(s);
            T.template reduce<1, true>(mx);
            double gsc = fmax(s[0] / (double)N, mx[0]);
            gsc = 1.0 / ((gsc > 1e-4) ? gsc : 1.0);
            gsc = fmin(fmax(gsc, 1e-4), 1e4);
            c *= gsc;
        }
        for (int i = tl; i < M; i += TL) {
            const double es = Es AT(i);
            for (int k = Jrb[i]; k < Jre[i]; ++k) Jsv AT(k) = es * Jv AT(k) * D AT(Jcol[k]);
            rl AT(i) *= es; ru AT(i) *= es;
        }
        for (int j = tl; j < N; j += TL) {
            const double dj = D AT(j);
            for (int k = Trb[j]; k < Tre[j]; ++k) Tsv AT(k) = dj * Tv AT(k) * Es AT(Tcol[k]);
            double dgv = c * pconst * dj * dj;
            if (useH)
                for (int k = Hrb[j]; k < Hre[j]; ++k) {
                    const double v = c * dj * Hv AT(k) * D AT(Hcol[k]);
                    Hsv AT(k) = v;
                    if (Hcol[k] == j) dgv += v;
                }
            hd AT(j) = dgv;
            q AT(j) *= c * dj;
            xl AT(j) /= dj;
            xu AT(j) /= dj;
        }
        __syncthreads();
    }

    // ================= interior point (ipm.cuh, lock step over the group) ===========================================
    // rows: s_u M_ZC, z_u M_YC, s_l M_RC, z_l M_BC, y M_I1, r_u M_AX / r_l M_I4, A x M_I3, J dx M_I2, w M_RW, t M_T, a M_YP / b M_TMP
    // cols: x N_X, s_u N_ZB, z_u N_YB, s_l N_RB, z_l N_KP, y N_MASK, r_u N_TMP2 / r_l N_I1, r_x N_R, rhs N_P, dx N_XT,
    //       lam_box N_TMP, diagonal of K N_DSH, a N_MINV / b N_XFIX
    bool solved = false, almost = false, infeasible = false, blowup = false;
    int iters = 0, nfact = 0;
    double rp_out = INFINITY, rd_out = INFINITY, rho_out = 0.0;
    {
        for (int j = tl; j < N; j += TL) {
            double v = (phase == SQPQP_PHASE_LP) ? xk[j] / NV(N_D) AT(j) : 0.0;
            v = fmin(fmax(v, NV(N_XL) AT(j)), NV(N_XU) AT(j));
            NV(N_X) AT(j) = v;
            NV(N_MASK) AT(j) = 0.0;
            NV(N_XT) AT(j) = 0.0;
        }
        for (int i = tl; i < M; i += TL) MV(M_I2) AT(i) = 0.0;
        __syncthreads();
        double cnt[1] = {0.0};
        ilv_rows<G>(T, M, lgJ, Jrb, Jre, Jcol, Jsv, NV(N_X), [&](int i, double ax) {
            const double rl_ = MV(M_RL) AT(i), ru_ = MV(M_RU) AT(i);
            const bool eq = rl_ == ru_;
            const bool uf = !eq && !isinf(ru_), lf = !eq && !isinf(rl_);
            const double su = uf ? fmax(ru_ - ax, 1.0) : 1.0, sl = lf ? fmax(ax - rl_, 1.0) : 1.0;
            MV(M_ZC) AT(i) = su; MV(M_YC) AT(i) = uf ? 1.0 : 0.0;
            MV(M_RC) AT(i) = sl; MV(M_BC) AT(i) = lf ? 1.0 : 0.0;
            MV(M_I1) AT(i) = 0.0;
            MV(M_I3) AT(i) = ax;
            MV(M_AX) AT(i) = eq ? ax - rl_ : (uf ? ax + su - ru_ : 0.0);
            MV(M_I4) AT(i) = lf ? -ax + sl + rl_ : 0.0;
            cnt[0] += (double)uf + (double)lf;
        });
        for (int j = tl; j < N; j += TL) {
            const double xl_ = NV(N_XL) AT(j), xu_ = NV(N_XU) AT(j), xj = NV(N_X) AT(j);
            const bool eq = xl_ == xu_;
            const bool uf = !eq && !isinf(xu_), lf = !eq && !isinf(xl_);
            const double su = uf ? fmax(xu_ - xj, 1.0) : 1.0, sl = lf ? fmax(xj - xl_, 1.0) : 1.0;
            NV(N_ZB) AT(j) = su; NV(N_YB) AT(j) = uf ? 1.0 : 0.0;
            NV(N_RB) AT(j) = sl; NV(N_KP) AT(j) = lf ? 1.0 : 0.0;
            NV(N_TMP2) AT(j) = eq ? xj - xl_ : (uf ? xj + su - xu_ : 0.0);
            NV(N_I1) AT(j) = lf ? -xj + sl + xl_ : 0.0;
            cnt[0] += (double)uf + (double)lf;
        }
        T.template reduce<1, false>(cnt);
        const double nin = fmax(cnt[0], 1.0);

        double delta = o.ipm_delta0, rho_p = o.ipm_rho0, rho_last = 0.0;
        int acc_cnt = 0;
        double rp_ref = INFINITY;
        double mu_t = o.ipm_mu0;
        double alpha = 0.0, sig_prev = 0.0, del_prev = delta;
        bool live = valid, failed = false;   // failed: left to the ADMM launch
        int it = 0;
        while (true) {
            if (it >= o.ipm_max_iter) {  // iteration cap (uniform): acceptable iterates are "solved to acceptable level"
                if (live && acc_cnt > 0) almost = true;
                break;
            }
            if (!__syncthreads_or(live)) break;
            double mx[8] = {0, 0, 0, 0, 0, 0, 0.0, -INFINITY};
            double m2[2] = {0.0, 0.0};
            double sums[2] = {0.0, 0.0};
            // ---- P1 rows: apply the previous step, multiplier, residual norms, weights, rhs coefficients ------------
            for (int i = tl; i < M; i += TL) {
                const double rl_ = MV(M_RL) AT(i), ru_ = MV(M_RU) AT(i), es = MV(M_ES) AT(i);
                double su = MV(M_ZC) AT(i), zu = MV(M_YC) AT(i), sl = MV(M_RC) AT(i), zl = MV(M_BC) AT(i);
                double yy = MV(M_I1) AT(i), rU = MV(M_AX) AT(i), rL = MV(M_I4) AT(i), ax = MV(M_I3) AT(i);
                const double jd = MV(M_I2) AT(i);
                const bool eq = rl_ == ru_, uf = !eq && !isinf(ru_), lf = !eq && !isinf(rl_);
                ax += alpha * jd;
                if (eq) { yy += alpha * (jd + rU) / del_prev; rU += alpha * jd; }
                if (uf) {
                    const SideDir d = side_dir(sig_prev - su * zu, zu, su, rU, jd, del_prev);
                    su += alpha * d.ds; zu += alpha * d.dz; rU += alpha * (jd + d.ds);
                }
                if (lf) {
                    const SideDir d = side_dir(sig_prev - sl * zl, zl, sl, rL, -jd, del_prev);
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
                if (live) {
                    MV(M_ZC) AT(i) = su; MV(M_YC) AT(i) = zu; MV(M_RC) AT(i) = sl; MV(M_BC) AT(i) = zl;
                    MV(M_I1) AT(i) = yy; MV(M_AX) AT(i) = rU; MV(M_I4) AT(i) = rL; MV(M_I3) AT(i) = ax;
                    MV(M_T) AT(i) = lam; MV(M_RW) AT(i) = wi; MV(M_YP) AT(i) = ai; MV(M_TMP) AT(i) = bi;
                }
            }
            // ---- P1 cols -----------------------------------------------------------------------------------------------
            for (int j = tl; j < N; j += TL) {
                const double xl_ = NV(N_XL) AT(j), xu_ = NV(N_XU) AT(j), Dj = NV(N_D) AT(j);
                double su = NV(N_ZB) AT(j), zu = NV(N_YB) AT(j), sl = NV(N_RB) AT(j), zl = NV(N_KP) AT(j);
                double yy = NV(N_MASK) AT(j), rU = NV(N_TMP2) AT(j), rL = NV(N_I1) AT(j), xj = NV(N_X) AT(j);
                const double dj = NV(N_XT) AT(j);
                const bool eq = xl_ == xu_, uf = !eq && !isinf(xu_), lf = !eq && !isinf(xl_);
                xj += alpha * dj;
                if (eq) { yy += alpha * (dj + rU) / del_prev; rU += alpha * dj; }
                if (uf) {
                    const SideDir d = side_dir(sig_prev - su * zu, zu, su, rU, dj, del_prev);
                    su += alpha * d.ds; zu += alpha * d.dz; rU += alpha * (dj + d.ds);
                }
                if (lf) {
                    const SideDir d = side_dir(sig_prev - sl * zl, zl, sl, rL, -dj, del_prev);
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
                if (live) {
                    NV(N_ZB) AT(j) = su; NV(N_YB) AT(j) = zu; NV(N_RB) AT(j) = sl; NV(N_KP) AT(j) = zl;
                    NV(N_MASK) AT(j) = yy; NV(N_TMP2) AT(j) = rU; NV(N_I1) AT(j) = rL; NV(N_X) AT(j) = xj;
                    NV(N_TMP) AT(j) = lamb; NV(N_DSH) AT(j) = wj + (useH ? 0.0 : NV(N_HD) AT(j));
                    NV(N_MINV) AT(j) = aj; NV(N_XFIX) AT(j) = bj;
                }
            }
            __syncthreads();
            // ---- P2: stationarity residual ---------------------------------------------------------------------------
            ilv_rows2<G>(T, N, lgT, Hrb, Hre, Hcol, Hsv, NV(N_X), useH, Trb, Tre, Tcol, Tsv, MV(M_T), [&](int j, double px, double aty) {
                const double lamb = NV(N_TMP) AT(j), qj = NV(N_Q) AT(j), id = 1.0 / NV(N_D) AT(j);
                if (!useH) px = NV(N_HD) AT(j) * NV(N_X) AT(j);
                const double r = px + qj + aty + lamb;
                if (live) NV(N_R) AT(j) = r;
                mx[4] = fmax(mx[4], fabs(aty + lamb) * id);
                m2[0] = fmax(m2[0], fabs(r));
                mx[1] = fmax(mx[1], fabs(r) * id);
                mx[3] = fmax(mx[3], fmax(fabs(px), fmax(fabs(aty), fabs(qj))) * id);
            });
            T.template reduce<8, true>(mx);
            T.template reduce<2, true>(m2);
            T.template reduce<2, false>(sums);
            // ---- per-instance tests (predicates, no branch around a barrier) -------------------------------------------
            if (live) {
                iters = it;
                const double ymx = m2[1], sup = sums[0];
                const double mu = sums[1] / nin;
                rp_out = mx[0];
                rd_out = mx[1] / c;
                const double scale_p = fmax(1.0, mx[2]), scale_d = fmax(1.0, mx[3] / c);
                const double comp_u = mx[6] / c, sc = fmax(1.0, ymx / c / 100.0);
                const double eps_c = (phase == SQPQP_PHASE_LP) ? 1e-3 * o.ipm_eps : o.ipm_eps;
                if (ymx > 1e12) {
                    blowup = true; almost = false; live = false; failed = true;
                } else if (mx[5] / c > 1e4 && mx[4] <= o.eps_inf * mx[5] && sup <= -o.eps_inf * mx[5]) {
                    infeasible = true; live = false;
                } else if (it >= 20 && it % 10 == 0 && rp_out > 1e4 * o.ipm_eps * scale_p && fabs(rp_out - rp_ref) <= 1e-3 * rp_out &&
                           rd_out <= 1e-5 * scale_d && mx[5] / c > 1e3) {
                    infeasible = true; live = false;
                } else if (rp_out <= o.ipm_eps * scale_p && rd_out <= o.ipm_eps * scale_d && comp_u <= eps_c * sc) {
                    solved = true; live = false;
                } else {
                    if (it % 10 == 0) rp_ref = rp_out;
                    const double acc_eps = 100.0 * o.ipm_eps;
                    const bool acceptable = rp_out <= acc_eps * scale_p && rd_out <= acc_eps * scale_d && comp_u <= 100.0 * eps_c * sc;
                    acc_cnt = acceptable ? acc_cnt + 1 : 0;
                    almost = rp_out <= 1e-6 * scale_p && rd_out <= 1e-6 * scale_d && comp_u <= 1e-6 * sc;
                    if (acc_cnt >= 8) {
                        almost = true; live = false;
                    } else if (!(mu == mu) || !(rd_out == rd_out)) {
                        almost = false; live = false; failed = true;
                    } else {
                        for (int gg = 0; gg < 60; ++gg) {  // monotone barrier update
                            const double comp = fmax(fabs(mx[6] - mu_t), fabs(-mx[7] - mu_t));
                            const double e_mu = fmax(m2[0], comp);
                            if (e_mu <= o.ipm_kappa_eps * mu_t && mu_t > o.ipm_mu_min) mu_t = fmax(o.ipm_mu_min, fmin(0.2 * mu_t, mu_t * sqrt(mu_t)));
                            else break;
                        }
                    }
                }
            }
            if (!__syncthreads_or(live)) break;
            // ---- assembly + factorisation with inertia correction (scalar shift rho_p), per instance -----------------------
            {
                bool need = live;
                int tries = 0;
                while (__syncthreads_or(need)) {
                    ilv_assemble<G>(T, C, W, useH ? Hsv : (const double*)nullptr, NV(N_DSH), rho_p, MV(M_RW), Jsv, need);
                    const bool ok = ilv_factor<G>(T, C, W, need);
                    if (need) {
                        ++nfact;
                        ++tries;
                        if (ok) need = false;
                        else {
                            rho_p = fmax(fmax(o.ipm_ic_growth * rho_p, rho_last > 0.0 ? rho_last / o.ipm_ic_decay : 1e-4), 1e-6);
                            if (rho_p > 1e8 || tries >= 30) { need = false; live = false; failed = true; almost = false; }
                        }
                    }
                }
                if (live) {
                    if (rho_p > 10.0 * o.ipm_rho0) rho_last = rho_p;
                    rho_out = rho_p;
                }
            }
            // ---- P3 / P4: right-hand side, Newton solve ---------------------------------------------------------------
            const double sigma_mu = mu_t;
            const double tau_k = fmax(o.ipm_tau, 1.0 - mu_t);
            for (int i = tl; i < M; i += TL)
                if (live) MV(M_T) AT(i) = fma(sigma_mu, MV(M_YP) AT(i), MV(M_TMP) AT(i));
            __syncthreads();
            ilv_rows<G>(T, N, lgT, Trb, Tre, Tcol, Tsv, MV(M_T), [&](int j, double tt) {
                const double tb = fma(sigma_mu, NV(N_MINV) AT(j), NV(N_XFIX) AT(j));
                if (live) NV(N_P) AT(j) = -NV(N_R) AT(j) - tt - tb;
            });
            __syncthreads();
            ilv_solve<G>(T, C, W, NV(N_P), NV(N_XT), live);
            // ---- P5: J dx and the step-to-boundary ratio ---------------------------------------------------------------
            double ratio[1] = {0.0};
            ilv_rows<G>(T, M, lgJ, Jrb, Jre, Jcol, Jsv, NV(N_XT), [&](int i, double jd) {
                const double rl_ = MV(M_RL) AT(i), ru_ = MV(M_RU) AT(i);
                const double su = MV(M_ZC) AT(i), zu = MV(M_YC) AT(i), sl = MV(M_RC) AT(i), zl = MV(M_BC) AT(i);
                const double rU = MV(M_AX) AT(i), rL = MV(M_I4) AT(i);
                if (live) MV(M_I2) AT(i) = jd;
                const bool eq = rl_ == ru_;
                if (!eq && !isinf(ru_)) {
                    const SideDir d = side_dir(sigma_mu - su * zu, zu, su, rU, jd, delta);
                    ratio[0] = fmax(ratio[0], fmax(step_ratio(su, d.ds), step_ratio(zu, d.dz)));
                }
                if (!eq && !isinf(rl_)) {
                    const SideDir d = side_dir(sigma_mu - sl * zl, zl, sl, rL, -jd, delta);
                    ratio[0] = fmax(ratio[0], fmax(step_ratio(sl, d.ds), step_ratio(zl, d.dz)));
                }
            });
            for (int j = tl; j < N; j += TL) {
                const double xl_ = NV(N_XL) AT(j), xu_ = NV(N_XU) AT(j);
                const double su = NV(N_ZB) AT(j), zu = NV(N_YB) AT(j), sl = NV(N_RB) AT(j), zl = NV(N_KP) AT(j);
                const double rU = NV(N_TMP2) AT(j), rL = NV(N_I1) AT(j), dj = NV(N_XT) AT(j);
                const bool eq = xl_ == xu_;
                if (!eq && !isinf(xu_)) {
                    const SideDir d = side_dir(sigma_mu - su * zu, zu, su, rU, dj, delta);
                    ratio[0] = fmax(ratio[0], fmax(step_ratio(su, d.ds), step_ratio(zu, d.dz)));
                }
                if (!eq && !isinf(xl_)) {
                    const SideDir d = side_dir(sigma_mu - sl * zl, zl, sl, rL, -dj, delta);
                    ratio[0] = fmax(ratio[0], fmax(step_ratio(sl, d.ds), step_ratio(zl, d.dz)));
                }
            }
            T.template reduce<1, true>(ratio);
            if (live) {
                alpha = 1.0;
                if (ratio[0] > 0.0) alpha = fmin(1.0, tau_k / ratio[0]);
                sig_prev = sigma_mu;
                del_prev = delta;
                delta = fmax(o.ipm_delta_min, delta * 0.3);
                if (rho_p > o.ipm_rho0) rho_p = fmax(o.ipm_rho0, rho_p / o.ipm_ic_decay);
                iters = it + 1;
            }
            ++it;
        }
        __syncthreads();
        (void)failed;
    }

    // ================= outputs (collect_solution!, subproblem_JuMP.jl:514-563) =======================================
    const bool ipm_done = infeasible || solved || almost;
    int status = SQPQP_MOI_NUMERICAL_ERROR;
    if (infeasible) status = SQPQP_MOI_LOCALLY_INFEASIBLE;
    else if (solved) status = SQPQP_MOI_LOCALLY_SOLVED;
    else if (almost) status = SQPQP_MOI_ALMOST_LOCALLY_SOLVED;
    const bool hand_over = valid && !ipm_done && o.method != 2;  // flagged for the ADMM launch (MODE 2 of k_solve_cta)
    if (valid && tl == 0) P.fb_flag[inst] = hand_over ? (blowup ? 2 : 1) : 0;
    const bool okst = solved || almost;
    double obj[1] = {0.0};
    {
        const double *x = NV(N_X), *q = NV(N_Q), *hd = NV(N_HD);
        if (useH)
            ilv_rows<G>(T, N, min(P.lgH, MAXLG), Hrb, Hre, Hcol, Hsv, x, [&](int r, double d) { obj[0] += x AT(r) * (0.5 * d + q AT(r)); });
        else
            for (int r = tl; r < N; r += TL) obj[0] += x AT(r) * (0.5 * hd AT(r) * x AT(r) + q AT(r));
        T.template reduce<1, false>(obj);
        obj[0] = okst ? obj[0] / c : 0.0;
    }
    if (valid && !hand_over) {
        const double *x = NV(N_X), *D = NV(N_D), *Es = MV(M_ES);
        // multipliers in the OSQP sign: rows y (eq) or z_u - z_l, same for the box
        double* op = P.o_p + (size_t)inst * n;
        double* ol = P.o_lam + (size_t)inst * m;
        double* oL = P.o_mxL + (size_t)inst * n;
        double* oU = P.o_mxU + (size_t)inst * n;
        double* os = P.o_slack + (size_t)inst * (P.S > 0 ? P.S : 1);
        for (int j = tl; j < n; j += TL) {
            double pv = 0.0, rcost = 0.0;
            if (okst) {
                const double yb = (NV(N_XL) AT(j) == NV(N_XU) AT(j)) ? NV(N_MASK) AT(j) : (NV(N_YB) AT(j) - NV(N_KP) AT(j));
                pv = D AT(j) * ((NV(N_XL) AT(j) == NV(N_XU) AT(j)) ? NV(N_XL) AT(j) : x AT(j));  // fixed column: its bound, exactly
                rcost = -yb / (D AT(j) * c);
            }
            op[j] = pv;
            oL[j] = rcost > 0.0 ? rcost : 0.0;
            oU[j] = rcost < 0.0 ? rcost : 0.0;
        }
        for (int i = tl; i < m; i += TL) {
            double v = 0.0;
            if (okst) {
                const double yc = (MV(M_RL) AT(i) == MV(M_RU) AT(i)) ? MV(M_I1) AT(i) : (MV(M_YC) AT(i) - MV(M_BC) AT(i));
                v = -(Es AT(i) * yc) / c;
            }
            ol[i] = v;
        }
        for (int s = tl; s < P.S; s += TL)
            os[s] = (okst && fr) ? D AT(n + s) * ((NV(N_XL) AT(n + s) == NV(N_XU) AT(n + s)) ? NV(N_XL) AT(n + s) : x AT(n + s)) : 0.0;
        if (tl == 0) {
            sqpqp_info& inf = P.o_info[inst];
            inf.moi_status = status;
            inf.admm_iters = 0; inf.cg_iters = 0; inf.polish_tries = 0; inf.polish_cg_iters = 0; inf.polished = 0;
            inf.rho_updates = 0; inf.checks = 0;
            inf.rho = 0.0;
            inf.rho_box_floor = okst ? rho_out : 0.0;
            inf.res_prim = okst ? rp_out : INFINITY;
            inf.res_dual = okst ? rd_out : INFINITY;
            inf.objective = obj[0];
            inf.ipm_iters = iters;
            inf.chol_factorizations = nfact;
        }
    } else if (valid && tl == 0) {  // handed to the ADMM launch: keep the interior-point statistics for it
        P.o_info[inst].ipm_iters = iters;
        P.o_info[inst].chol_factorizations = nfact;
    }
    __syncthreads();
#undef NV
#undef MV
#undef AT
}

// One CTA per group of G instances; NT threads = NT / G task lanes.
template <int G, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_solve_ilv(const __grid_constant__ Prob P, const __grid_constant__ DevOpts O, int phase,
                                                       const __grid_constant__ IlvDev X) {
    __shared__ double sh[2 * ILV_KMAX * (NT / 32) * G];
    extern __shared__ double dsm[];
    for (int grp = blockIdx.x; grp < X.ngroups; grp += gridDim.x) {
        if (P.active) {  // skip a group with no active instance (uniform)
            bool any = false;
            for (int k = 0; k < G; ++k) {
                const int b = grp * G + k;
                any = any || (b < P.batch && P.active[b]);
            }
            if (!any) continue;
        }
        IlvTeam<G> T(sh);
        ilv_solve_group<G>(T, P, O.o, X, grp, phase, dsm);
        __syncthreads();
    }
}
