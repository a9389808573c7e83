// spmv.cuh -- batched CSR sparse matrix-vector products with one shared pattern:
//      y_b = J_b x_b   (Jacobian * p, sqp_trust_region.jl:343,492)
//      y_b = J_b' x_b  (Jac' * lambda, common.jl:17)
//      y_b = H_b x_b   (Hessian * p, sqp_trust_region.jl:490)
//
// The ACOPF matrices have very short rows (J: 1-12 entries, H: 1-25), so a warp- or even a
// sub-warp-per-row kernel leaves most lanes idle and reads the value stream in fragments.  This is
// the CSR-stream scheme instead: the rows are cut ONCE (host, shared by the batch) into blocks of
// at most SPMV_CHUNK value slots; a CTA streams the value slots of its block with fully coalesced
// loads (thread k -> slot k), multiplies by the gathered x (L1/L2 resident) into shared memory, and
// then sums each row's products from shared memory in ascending slot order -- deterministic, and
// the same summation order as a sequential CSR row loop.  HBM traffic per instance is the fp64
// value stream + x + y; the int32 pattern is shared by the batch and stays in L2.
//
// Rows longer than a chunk get a block of their own and are reduced chunk by chunk.
#pragma once
#include "team.cuh"

#define SPMV_CHUNK 2048
#define SPMV_THREADS 256
#define SPMV_INST 1  // instances a CTA handles per trip (2 was measured slower: 2.0-2.7 TB/s against 2.5-3.1)

struct SpmvPlan {
    int nblocks;
    const int* blk_row;   // [nblocks+1] first row of each block
    const int* rb;        // row begin
    const int* re;        // row end (may stop before the next row's begin: J keeps its slack slots at the row tail)
    const int* col;
    int ncols;            // entries with col >= ncols are skipped (slack columns in the normal phase)
    int nrows, nnz;       // rows of y; value slots per instance
};

__global__ void __launch_bounds__(SPMV_THREADS) k_spmv_stream(SpmvPlan S, const double* __restrict__ vals, const double* __restrict__ x,
                                                              double* __restrict__ y, int xstride, int batch) {
    __shared__ double prod[SPMV_INST][SPMV_CHUNK];
    const int blk = blockIdx.x;
    const int r0 = S.blk_row[blk], r1 = S.blk_row[blk + 1];
    const int s0 = S.rb[r0], s1 = S.re[r1 - 1];
    for (int b0 = blockIdx.y * SPMV_INST; b0 < batch; b0 += gridDim.y * SPMV_INST) {
        if (s1 - s0 <= SPMV_CHUNK) {
            // all value and column loads of the thread's slots first (independent, coalesced), then the gathers
            constexpr int PER = SPMV_CHUNK / SPMV_THREADS;
            int cc[PER];
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                const int k = s0 + threadIdx.x + i * SPMV_THREADS;
                cc[i] = k < s1 ? S.col[k] : S.ncols;
            }
#pragma unroll
            for (int u = 0; u < SPMV_INST; ++u) {
                const int b = b0 + u;
                if (b >= batch) break;
                const double* __restrict__ v = vals + (size_t)b * S.nnz;
                const double* __restrict__ xb = x + (size_t)b * xstride;
                double vv[PER];
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    const int k = s0 + threadIdx.x + i * SPMV_THREADS;
                    vv[i] = cc[i] < S.ncols ? v[k] : 0.0;
                }
#pragma unroll
                for (int i = 0; i < PER; ++i) prod[u][threadIdx.x + i * SPMV_THREADS] = cc[i] < S.ncols ? vv[i] * xb[cc[i]] : 0.0;
            }
            __syncthreads();
            for (int u = 0; u < SPMV_INST; ++u) {
                const int b = b0 + u;
                if (b >= batch) break;
                double* __restrict__ yb = y + (size_t)b * S.nrows;
                for (int r = r0 + threadIdx.x; r < r1; r += SPMV_THREADS) {
                    double acc = 0.0;
                    for (int k = S.rb[r] - s0, e = S.re[r] - s0; k < e; ++k) acc += prod[u][k];
                    yb[r] = acc;
                }
            }
            __syncthreads();
        } else {  // one long row: chunked block reduction in a fixed order
            for (int u = 0; u < SPMV_INST; ++u) {
                const int b = b0 + u;
                if (b >= batch) break;
                const double* __restrict__ v = vals + (size_t)b * S.nnz;
                const double* __restrict__ xb = x + (size_t)b * xstride;
                double acc = 0.0;
                for (int k = s0 + threadIdx.x; k < s1; k += SPMV_THREADS) {
                    const int c = S.col[k];
                    if (c < S.ncols) acc = fma(v[k], xb[c], acc);
                }
                prod[0][threadIdx.x] = acc;
                __syncthreads();
                if (threadIdx.x == 0) {
                    double t = 0.0;
                    for (int i = 0; i < SPMV_THREADS; ++i) t += prod[0][i];
                    y[(size_t)b * S.nrows + r0] = t;
                }
                __syncthreads();
            }
        }
    }
}

// host: cut rows into blocks of <= SPMV_CHUNK value slots (span rb[r0] .. re[r1-1])
inline void spmv_blocks(int nrows, const int* rb, const int* re, std::vector<int>& blk_row) {
    blk_row.clear();
    blk_row.push_back(0);
    int r0 = 0;
    while (r0 < nrows) {
        int r1 = r0 + 1;
        while (r1 < nrows && re[r1] - rb[r0] <= SPMV_CHUNK && r1 - r0 < 4 * SPMV_CHUNK) ++r1;
        blk_row.push_back(r1);
        r0 = r1;
    }
}
