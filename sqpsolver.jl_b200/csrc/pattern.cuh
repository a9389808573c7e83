// pattern.cuh -- K1: COO -> CSR pattern + ordered-duplicate permutation, built on device;
//                K2: per-iterate value scatter (deterministic segmented sum, no atomics).
//
// Replaces  sparse(j_row,j_col,ones,m,n) / sparse(h_row,h_col,ones,n,n)
//           (sqp_trust_region.jl:47-48, 56-57)               -> build_pattern()
//           fill!(nzval,0); A[r,c] += v  for k ascending      -> k_scatter
//           (sqp.jl:111-117; symmetric mirror sqp.jl:92-103)
//
// An "entry" e is (row, col, src): it adds values[src] to slot (row, col).  Entries of
// one slot are summed in ascending e, starting from 0.0, which is exactly Julia's
// accumulation order (e ascending <=> COO index k ascending).  The atomics used to
// bucket entries by row only decide an intermediate position; the per-row rank sort
// on the unique key (col, e) makes the final layout independent of their order.
#pragma once
#include "common.cuh"

// entry generators ----------------------------------------------------------------
// J ext: e < nnz: (r,c) from COO; e >= nnz: slack column e-nnz.  transpose swaps.
__global__ void k_entries_jac(int64_t nnz, const int64_t* __restrict__ r1, const int64_t* __restrict__ c1, int S,
                              const int* __restrict__ slack_row, int n, int transpose, int* erow, int* ecol, int* esrc) {
    int64_t L = nnz + S;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < L; e += (int64_t)gridDim.x * blockDim.x) {
        int r, c;
        if (e < nnz) {
            r = (int)(r1[e] - 1);
            c = (int)(c1[e] - 1);
        } else {
            r = slack_row[e - nnz];
            c = n + (int)(e - nnz);
        }
        erow[e] = transpose ? c : r;
        ecol[e] = transpose ? r : c;
        esrc[e] = (int)e;
    }
}
// H: e = 2k -> (r,c); e = 2k+1 -> (c,r) iff r != c (else invalid, row = -1); src = k.
__global__ void k_entries_hess(int64_t nnz, const int64_t* __restrict__ r1, const int64_t* __restrict__ c1, int* erow,
                               int* ecol, int* esrc) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
        int r = (int)(r1[k] - 1), c = (int)(c1[k] - 1);
        erow[2 * k] = r;
        ecol[2 * k] = c;
        esrc[2 * k] = (int)k;
        erow[2 * k + 1] = (r != c) ? c : -1;
        ecol[2 * k + 1] = r;
        esrc[2 * k + 1] = (int)k;
    }
}

__global__ void k_count_rows(int L, const int* __restrict__ erow, int* cnt) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < L; e += gridDim.x * blockDim.x)
        if (erow[e] >= 0) atomicAdd(&cnt[erow[e]], 1);
}

// single-block exclusive scan: out[0..n] (n+1 values), in[0..n-1]
__global__ void k_exscan(int n, const int* __restrict__ in, int* out) {
    __shared__ int sh[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        int i = base + threadIdx.x;
        int v = (i < n) ? in[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        int incl = sh[threadIdx.x];
        if (i < n) out[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

__global__ void k_bucket(int L, const int* __restrict__ erow, const int* __restrict__ ecol, const int* __restrict__ rstart,
                         int* cursor, int* tcol, int* tent) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < L; e += gridDim.x * blockDim.x) {
        int r = erow[e];
        if (r < 0) continue;
        int pos = rstart[r] + atomicAdd(&cursor[r], 1);
        tcol[pos] = ecol[e];
        tent[pos] = e;
    }
}

// one warp per row: rank sort on the unique key (col, entry id)
__global__ void k_sort_rows(int nrows, const int* __restrict__ rstart, const int* __restrict__ tcol,
                            const int* __restrict__ tent, int* scol, int* sent) {
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < nrows; r += nwarps) {
        int b = rstart[r], len = rstart[r + 1] - b;
        for (int i = lane; i < len; i += 32) {
            int ci = tcol[b + i], ei = tent[b + i], rank = 0;
            for (int j = 0; j < len; ++j) {
                int cj = tcol[b + j], ej = tent[b + j];
                rank += (cj < ci) || (cj == ci && ej < ei);
            }
            scol[b + rank] = ci;
            sent[b + rank] = ei;
        }
    }
}

__global__ void k_heads(int nrows, const int* __restrict__ rstart, const int* __restrict__ scol, int* head) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
        int b = rstart[r], e = rstart[r + 1];
        for (int p = b; p < e; ++p) head[p] = (p == b) || (scol[p] != scol[p - 1]);
    }
}

// slotof = exclusive scan of head (length Lv+1)
__global__ void k_emit(int nrows, int Lv, const int* __restrict__ rstart, const int* __restrict__ scol,
                       const int* __restrict__ sent, const int* __restrict__ esrc, const int* __restrict__ head,
                       const int* __restrict__ slotof, int* row_ptr, int* col_idx, int* seg_ptr, int* seg_src) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int r = tid; r <= nrows; r += nt) row_ptr[r] = slotof[rstart[r]];
    for (int p = tid; p < Lv; p += nt) {
        seg_src[p] = esrc[sent[p]];
        if (head[p]) {
            col_idx[slotof[p]] = scol[p];
            seg_ptr[slotof[p]] = p;
        }
    }
    if (tid == 0) seg_ptr[slotof[Lv]] = Lv;
}

// row end excluding columns >= ncols_normal (slack columns sort last)
__global__ void k_row_end_normal(int nrows, const int* __restrict__ row_ptr, const int* __restrict__ col, int ncn,
                                 int* re_n) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
        int e = row_ptr[r + 1];
        while (e > row_ptr[r] && col[e - 1] >= ncn) --e;
        re_n[r] = e;
    }
}

// K2: values of every slot of J ext, its transpose and H, for all instances.
//   src < nsrc : vals[b*nsrc + src]    else consts[src - nsrc]
struct ScatterJob {
    int nslots, nsrc;
    const int *seg_ptr, *seg_src;
    const double* vals;     // [batch][nsrc]
    const double* consts;   // shared
    double* out;            // [batch][nslots]
    double* out_ilv;        // nullable: the same values group-interleaved, [batch / G][nslots][G] (ilv.cuh)
    int G;
};
__global__ void k_scatter(ScatterJob a, ScatterJob b, ScatterJob c, int batch) {
    const ScatterJob* jobs[3] = {&a, &b, &c};
    for (int j = 0; j < 3; ++j) {
        const ScatterJob& J = *jobs[j];
        if (J.nslots == 0 || J.vals == nullptr) continue;
        int64_t total = (int64_t)J.nslots * batch;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            int inst = (int)(i / J.nslots), s = (int)(i - (int64_t)inst * J.nslots);
            const double* v = J.vals + (int64_t)inst * J.nsrc;
            double acc = 0.0;
            for (int p = J.seg_ptr[s]; p < J.seg_ptr[s + 1]; ++p) {
                int src = J.seg_src[p];
                acc += (src < J.nsrc) ? v[src] : J.consts[src - J.nsrc];
            }
            J.out[(int64_t)inst * J.nslots + s] = acc;
            if (J.out_ilv) J.out_ilv[((int64_t)(inst / J.G) * J.nslots + s) * J.G + (inst % J.G)] = acc;
        }
    }
}
