// team.cuh -- the "team" abstraction: the set of threads that solves ONE QP instance.
//
//   CtaTeam  : one thread block per instance; sync = __syncthreads (bar.sync).  Used for
//              the batched workload (1024 independent ACOPF instances -> 1024 CTAs, no
//              inter-CTA communication at all) and for small single instances.
//   GridTeam : the whole cooperative grid (148 SMs x resident CTAs) works on one instance;
//              sync = grid.sync().  Used for the ~2000-bus single instance where one SM's
//              L2 bandwidth would be the limiter.
//
// Both give the same device algorithm (admm.cuh) three primitives: sync(), an elementwise
// index mapping, and deterministic fused reductions (fixed summation order -> bit-reproducible
// across runs; warp shuffles, then shared memory, then -- GridTeam only -- a second pass over
// per-block partials in global memory; no atomics anywhere).
#pragma once
#include "common.cuh"

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

struct CtaTeam {
    double* sh;  // shared: [2][SQPQP_MAX_RED][32]
    int flip;
    __device__ CtaTeam(double* s) : sh(s), flip(0) {}
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int size() const { return blockDim.x; }
    __device__ __forceinline__ void sync() { __syncthreads(); }

    // fused reduction of K values; IS_MAX selects max instead of sum.  Result in v[] for all threads.
    template <int K, bool IS_MAX>
    __device__ void reduce(double (&v)[K]) {
        double* buf = sh + flip * (SQPQP_MAX_RED * 32);
        flip ^= 1;
        int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double r = IS_MAX ? warp_max(v[k]) : warp_sum(v[k]);
            if (lane == 0) buf[k * 32 + w] = r;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double r = (lane < nw) ? buf[k * 32 + lane] : (IS_MAX ? -INFINITY : 0.0);
            v[k] = IS_MAX ? warp_max(r) : warp_sum(r);
        }
        // the other buffer is used next time, so no trailing barrier is needed: a thread
        // can only reach the write of call i+2 after passing the barrier of call i+1,
        // which every thread reaches only after finishing the reads of call i.
    }
};

struct GridTeam {
    double* sh;
    double* g;  // global: [2][SQPQP_MAX_RED][stride]
    int stride;
    int flip;
    cg::grid_group grid;
    __device__ GridTeam(double* s, double* gbuf, int st) : sh(s), g(gbuf), stride(st), flip(0), grid(cg::this_grid()) {}
    __device__ __forceinline__ int tid() const { return blockIdx.x * blockDim.x + threadIdx.x; }
    __device__ __forceinline__ int size() const { return gridDim.x * blockDim.x; }
    __device__ __forceinline__ void sync() { grid.sync(); }

    template <int K, bool IS_MAX>
    __device__ void reduce(double (&v)[K]) {
        int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        double* gb = g + (size_t)flip * SQPQP_MAX_RED * stride;
        flip ^= 1;
        // block level
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double r = IS_MAX ? warp_max(v[k]) : warp_sum(v[k]);
            if (lane == 0) sh[k * 32 + w] = r;
        }
        __syncthreads();
        if (w == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double r = (lane < nw) ? sh[k * 32 + lane] : (IS_MAX ? -INFINITY : 0.0);
                r = IS_MAX ? warp_max(r) : warp_sum(r);
                if (lane == 0) gb[k * stride + blockIdx.x] = r;
            }
        }
        grid.sync();
        // every block re-reduces all partials in the same fixed order
        int nb = gridDim.x;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double r = IS_MAX ? -INFINITY : 0.0;
            for (int b = threadIdx.x; b < nb; b += blockDim.x) {
                double t = gb[k * stride + b];
                r = IS_MAX ? fmax(r, t) : r + t;
            }
            r = IS_MAX ? warp_max(r) : warp_sum(r);
            __syncthreads();  // sh reuse
            if (lane == 0) sh[w] = r;
            __syncthreads();
            double q = (lane < nw) ? sh[lane] : (IS_MAX ? -INFINITY : 0.0);
            v[k] = IS_MAX ? warp_max(q) : warp_sum(q);
        }
        __syncthreads();
    }
};

// ---- sparse row products ------------------------------------------------------------
// 2^lg lanes cooperate on one row (sub-warp per row); lg is chosen on the host from the
// row-length histogram of each matrix (short ACOPF rows -> 1..8 lanes; long rows -> 32).
// f(row, dot) is called by lane 0 of the sub-warp.
// Two rows per sub-warp and trip, walked together: the index -> value -> gather chains of both rows are in flight at once (a
// single row per trip pays the three dependent L2 round trips once per row; the summation order within a row is unchanged,
// so the results are bit-identical to the one-row loop).
template <class Team, class F>
__device__ __forceinline__ void csr_rows(Team& T, int nrows, int lg, const int* __restrict__ rb,
                                         const int* __restrict__ re, const int* __restrict__ col,
                                         const double* __restrict__ val, const double* __restrict__ x, F f) {
    const int L = 1 << lg, lane = T.tid() & (L - 1), sub = T.tid() >> lg, nsub = T.size() >> lg;
    for (int r0 = 0; r0 < nrows; r0 += 2 * nsub) {
        const int ra = r0 + sub, rc = ra + nsub;
        const bool ona = ra < nrows, onc = rc < nrows;
        int ka = ona ? rb[ra] + lane : 0, kc = onc ? rb[rc] + lane : 0;
        const int ea = ona ? re[ra] : 0, ec = onc ? re[rc] : 0;
        double acca = 0.0, accc = 0.0;
        while (ka < ea || kc < ec) {
            const bool oa = ka < ea, oc = kc < ec;
            const int ca = col[oa ? ka : 0], cc = col[oc ? kc : 0];
            const double va = val[oa ? ka : 0], vc = val[oc ? kc : 0];
            const double xa = x[ca], xc = x[cc];
            if (oa) acca = fma(va, xa, acca);
            if (oc) accc = fma(vc, xc, accc);
            ka += L; kc += L;
        }
        for (int o = L >> 1; o > 0; o >>= 1) {
            acca += __shfl_xor_sync(0xffffffffu, acca, o);
            accc += __shfl_xor_sync(0xffffffffu, accc, o);
        }
        if (ona && lane == 0) f(ra, acca);
        if (onc && lane == 0) f(rc, accc);
    }
}

// two CSR matrices with the same row count in one pass: dot1 = A1[r,:]*x1, dot2 = A2[r,:]*x2
template <class Team, class F>
__device__ __forceinline__ void csr_rows2(Team& T, int nrows, int lg, const int* __restrict__ rb1,
                                          const int* __restrict__ re1, const int* __restrict__ col1,
                                          const double* __restrict__ val1, const double* __restrict__ x1, bool use1,
                                          const int* __restrict__ rb2, const int* __restrict__ re2,
                                          const int* __restrict__ col2, const double* __restrict__ val2,
                                          const double* __restrict__ x2, F f) {
    const int L = 1 << lg, lane = T.tid() & (L - 1), sub = T.tid() >> lg, nsub = T.size() >> lg;
    for (int r0 = 0; r0 < nrows; r0 += nsub) {
        int r = r0 + sub;
        const bool on = r < nrows;
        // both products of the row walked together (their chains are independent)
        int k1 = (on && use1) ? rb1[r] + lane : 0, k2 = on ? rb2[r] + lane : 0;
        const int e1 = (on && use1) ? re1[r] : 0, e2 = on ? re2[r] : 0;
        double a1 = 0.0, a2 = 0.0;
        while (k1 < e1 || k2 < e2) {
            const bool o1 = k1 < e1, o2 = k2 < e2;
            const int c1 = col1[o1 ? k1 : 0], c2 = col2[o2 ? k2 : 0];
            const double v1 = val1[o1 ? k1 : 0], v2 = val2[o2 ? k2 : 0];
            const double y1 = x1[c1], y2 = x2[c2];
            if (o1) a1 = fma(v1, y1, a1);
            if (o2) a2 = fma(v2, y2, a2);
            k1 += L; k2 += L;
        }
        for (int o = L >> 1; o > 0; o >>= 1) {
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if (on && lane == 0) f(r, a1, a2);
    }
}

// row-wise max_k |val[k]| * s[col[k]]  (Ruiz equilibration)
template <class Team, class F>
__device__ __forceinline__ void csr_rows_absmax(Team& T, int nrows, const int* __restrict__ rb,
                                                const int* __restrict__ re, const int* __restrict__ col,
                                                const double* __restrict__ val, const double* __restrict__ s, F f) {
    for (int r = T.tid(); r < nrows; r += T.size()) {
        double a = 0.0;
        for (int k = rb[r]; k < re[r]; ++k) a = fmax(a, fabs(val[k]) * s[col[k]]);
        f(r, a);
    }
}
