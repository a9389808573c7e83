// common.cuh -- shared declarations of the sqpqp engine (sm_100a only).
//
// Data layout in HBM:
//   * sparsity patterns (int32 CSR) are built once on device and shared by all
//     instances of a batch;
//   * every per-instance array is instance-major: base + inst * stride, so one
//     team (a CTA, or the whole cooperative grid) streams a contiguous block;
//   * J is stored as CSR of [J | S] (S = slack columns of the feasibility-
//     restoration LP, sorted last in each row) with TWO row-end arrays, so the
//     normal phase and the FR phase share one value array; its transpose is
//     stored as CSR too (rows 0..n-1 of it are exactly Julia's CSC of J).
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <math.h>
#include "../../include/sqpqp.h"

namespace cg = cooperative_groups;

#define SQPQP_MAX_RED 8  // values per fused reduction

// ---- vector slots ----------------------------------------------------------------
// N-type vectors have per-instance stride Ne = n + S, M-type vectors stride m.
enum NVec {
    N_Q = 0, N_XL, N_XU, N_D, N_X, N_ZB, N_YB, N_RB, N_XT, N_R, N_P, N_KP, N_MINV, N_DSH, N_XFIX, N_MASK,
    N_XW, N_YBW, N_HD, N_TMP, N_TMP2, N_I1, N_COUNT
};
enum MVec {
    M_RL = 0, M_RU, M_ES, M_ZC, M_YC, M_RC, M_T, M_RW, M_BC, M_YP, M_YCW, M_TMP, M_AX, M_I1, M_I2, M_I3, M_I4, M_COUNT
};

struct Csr {
    const int* rb;   // row begin [nrows]  (== row_ptr)
    const int* re;   // row end   [nrows]  (row_ptr+1, or the "normal phase" end)
    const int* col;  // [nnz]
};

// device view of the symbolic Cholesky analysis (symbolic.hpp); all arrays shared by the batch
struct CholDev {
    int n, nnzL, nlev;
    const int *perm;
    const int *Lp, *Li;
    const int *Rp, *Rc, *Ri;
    const int *lev_ptr, *lev_cols;
    const int *fd_ptr, *fo_ptr, *f_ent, *fp_ptr, *fp_a, *fp_b, *ent_diag;
    const int *as_ptr, *as_a, *as_b, *as_r, *as_h, *as_d;
};

struct Prob {
    int n, m, mlin, S, Ne, batch;
    int nnzJ, nnzT, nnzH;         // slots per instance (J ext, its transpose, H symmetric)
    int has_hess;
    // shared patterns
    const int *J_rb, *J_re_n, *J_re_e, *J_col;  // [m]: CSR of [J|S]; re_n excludes slack columns
    const int *T_rb, *T_col;                    // [Ne+1]: CSR of [J|S]^T (row_ptr form)
    const int *H_rb, *H_col;                    // [n+1]
    const int* slack_row;                       // [S] row of each slack column
    const double* slack_sign;                   // [S] +1 / -1
    int lgJn, lgJe, lgT, lgH;                   // log2(lanes per row) per matrix
    // per-instance values (unscaled, scaled)
    double *Jv, *Tv, *Hv, *Jsv, *Tsv, *Hsv;
    // per-instance NLP data
    const double *df, *E, *Eov;  // Eov: SOC override or nullptr
    const double *gL, *gU, *xL, *xU;
    int gstride, xstride;        // 0 when bounds are shared by all instances
    const double *xk, *delta;
    const int* active;           // nullable
    // workspace
    double* nv[N_COUNT];
    double* mv[M_COUNT];
    signed char *codeC, *codeB, *prevC, *prevB, *triedC, *triedB;  // active-set codes
    double* rho_w;               // [batch] warm-start rho (0 = none)
    // outputs
    double *o_p, *o_lam, *o_mxL, *o_mxU, *o_slack;
    sqpqp_info* o_info;
    // interior-point path: symbolic Cholesky (null n = unavailable), per-instance factor values
    // [batch][nnzL] and permuted solve scratch [batch][n]
    CholDev chol;      // QP / SOC / LP-projection phases: n columns, P = H or 2I
    CholDev chol_fr;   // feasibility-restoration LP: n + S columns ([J|S]), P = 0
    int has_chol, has_chol_fr;
    double *Lval, *yw, *Lval_fr, *yw_fr;
    // grid-team reduction scratch: [2][SQPQP_MAX_RED][maxblocks]
    double* gred;
    int gred_stride;
};

// Where each per-instance scratch array of the solve lives for a CtaTeam: offset (in doubles)
// into the CTA's dynamic shared memory, or -1 = stays in global memory (L2).  Filled greedily
// on the host in order of how hot the array is in the PCG loop (see place_arrays()).
struct Placement {
    int n_off[N_COUNT];
    int m_off[M_COUNT];
    int jsv, tsv, hsv;
    int lval, yw;
    int total;  // doubles
};

#define CUDA_OK(call)                                                       \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) return fail_cuda(h, e__, #call, __LINE__);   \
    } while (0)
