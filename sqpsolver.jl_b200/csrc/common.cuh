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
#define GD_NB 32         // panel width of the grid team's blocked dense tail (chol.cuh)
#define GD_LD 33         // padded leading dimension of its 32 x 32 shared-memory tiles
#define RING_S 3         // stages of the shared-memory ring the index programs are streamed through (chol.cuh)

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

// device view of the symbolic Cholesky analysis (symbolic.hpp); all arrays shared by the batch.
// Columns are numbered level-major: sparse level l = columns lev_ptr[l] .. lev_ptr[l+1]-1, then the
// dense tail n0 .. n-1 (T columns, factorised as a packed dense matrix in shared memory).
struct CholDev {
    int n, nnzL, nlev, n0, T;
    int Tpad;               // T rounded up to the panel width of the dense code (4: CTA team, 32: grid team); padding rows are identity
    int nphase, n_aslot, nslotJ;
    int fused_fwd;          // the factor program carries the forward sweep of the Newton solve (symbolic.hpp: fuse_fwd)
    const int *perm;
    const int *Lp, *Li;
    const int *Rp, *Rmid;
    const int2 *Rci;        // (value index, column) of the strictly-lower CSR
    const int *lev_ptr;
    const int2 *fp_ab;      // pairs of value indices
    const int4 *ftask;      // factorisation tasks  (entry | has_K << 30, first pair, end pair, aux)
    const int4 *fphase;     // factorisation phases (first task, end task, max pairs, kind)
    const int4 *aslot;      // assembly slots (entry | lg << 26 | leader << 29, first term of the lane, end term, P index | -1)
    const int *aslot_d;     // per assembly slot: original column for the diagonal term d[], or -1
    const int2 *as_ab;      // assembly terms (wJ index, Jv index)
    const int *jrow;        // row of every J value slot
    // ring programs (symbolic.hpp: build_ring_program) for the resident CTA team; ring_ok = 0: not built
    int ring_ok, ring_nL, ring_stage_words;
    const int *rprog;                 // chunk images, 16-byte aligned
    int rseg_n[3];                    // chunks per segment (0 assembly + factor, 1 forward + tail rhs, 2 backward)
    int rseg_off[3][RING_S];          // word offset of the first RING_S chunks of each segment (-1: none)
    int rseg_bytes[3][RING_S];
};

// per-instance numeric state of the factorisation, resolved for one team
struct CholWork {
    double* L;      // [nnzL] factor values (global / L2; tail entries hold the assembled K only)
    double* D;      // [T(T+1)/2] dense tail, packed row-major lower triangle (shared memory)
    double* col;    // [T] scaled pivot column of the dense factorisation (shared memory)
    double* dinv;   // [n] 1 / L_jj
    double* yw;     // [n] triangular-solve scratch in permuted order
    double* wJ;     // [nslotJ] w[row] * Jv per J value slot (global)
    double* gsm;    // grid team: per-CTA shared scratch of the blocked dense code (3 x 32 x 33 doubles), else null
    int oL, oyw, odinv, oD;  // ring mode: the same four arrays as offsets (doubles) into the CTA's dynamic shared memory
};

// Interior-point state of an instance the throughput launch handed over after its iteration quota (launch_solve: hand-off):
// the scalars of the loop; the iterate itself (x, slacks, multipliers, tracked residuals, last direction) is in the
// per-instance work vectors in global memory.
struct IpmState {
    double delta, rho_p, rho_last, mu_t, alpha, sig_prev, del_prev, rp_ref, nin;
    int it, nfact, acc_cnt, pad;
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
    int* fb_flag;                // [batch] 0 solved by the interior-point launch, 1 needs the ADMM launch, 2 ditto after a blow-up,
                                 //         3 handed over to the resident launch after its iteration quota (state in ipm_state)
    IpmState* ipm_state;         // [batch]
    // interior-point path: symbolic Cholesky (null n = unavailable), per-instance factor values
    // [batch][nnzL] and permuted solve scratch [batch][n]
    CholDev chol;      // QP / SOC / LP-projection phases: n columns, P = H or 2I
    CholDev chol_fr;   // feasibility-restoration LP: n + S columns ([J|S]), P = 0
    int has_chol, has_chol_fr;
    double *Lval, *yw, *Lval_fr, *yw_fr;
    double *wJ;               // [batch][nnzJ]
    double *dinv, *dinv_fr;   // [batch][n], [batch][Ne] (used when the shared-memory budget cannot hold them)
    double *Dtail, *Dtail_fr; // grid team (one large instance): packed dense tail of the factor in global memory (L2-resident)
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
    int dtail, dcol, dinv;  // dense tail of the factor, its pivot column, inverse diagonal (interior-point path)
    int ring;               // RING_S stages of chol.ring_stage_words words for the streamed index programs, or -1 (slot lists from L2)
    int vec_resident;       // 1 if any work vector / matrix value array is placed (else only the factorisation parts)
    int total;  // doubles
};

// state of the shared-memory ring the index programs are streamed through (chol.cuh); uniform across the CTA
struct Ring {
    const int* prog;       // global: chunk images
    int stage0w;           // first stage as a WORD offset into the CTA's dynamic shared memory (RING_S x stage_words words)
    int stage_words;
    unsigned bar0;         // shared address of mbarrier 0 (8 bytes each)
    unsigned phase_bits;   // expected parity per stage
    int head;              // chunks of the current segment consumed so far
    int inflight;          // chunks issued, not yet consumed
    int seg;               // segment primed or in progress (-1: none)
    bool on;
};

// ---- optional in-kernel phase profile (build with -DSQPQP_PROF; tools/gpu_prof.py) ----------------
// Thread 0 of every CTA adds the clock64 cycles it spent in each segment of the solve to a global
// table; the sum over CTAs gives the share of CTA-time per segment.  Compiled out by default.
enum ProfSeg {
    PS_PROLOGUE = 0, PS_RESID, PS_WEIGHTS, PS_ASSEMBLE, PS_FACTOR_SPARSE, PS_ASSEMBLE_SLOTS, PS_FACTOR_DENSE, PS_RHS, PS_FWD, PS_TAIL,
    PS_BWD, PS_RATIO, PS_UPDATE, PS_EPILOGUE, PS_OTHER, PS_COUNT,
    // ring detail (thread 0; included in the segments above): waiting for a chunk, working on it, at its barrier, chunks
    PS_RING_WAIT = 16, PS_RING_WORK, PS_RING_BAR, PS_RING_CHUNKS
};
#ifdef SQPQP_PROF
__device__ unsigned long long g_prof[32];
struct Prof {
    long long t0;
    __device__ __forceinline__ void start() { t0 = clock64(); }
    __device__ __forceinline__ void lap(int k) {
        long long t = clock64();
        if (threadIdx.x == 0) atomicAdd(&g_prof[k], (unsigned long long)(t - t0));
        t0 = t;
    }
    __device__ __forceinline__ void add(int k, long long v) {
        if (threadIdx.x == 0) atomicAdd(&g_prof[k], (unsigned long long)v);
    }
};
#define SQPQP_CLK() clock64()
#else
struct Prof {
    __device__ __forceinline__ void start() {}
    __device__ __forceinline__ void lap(int) {}
    __device__ __forceinline__ void add(int, long long) {}
};
#define SQPQP_CLK() 0ll
#endif

#define CUDA_OK(call)                                                       \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) return fail_cuda(h, e__, #call, __LINE__);   \
    } while (0)
