/*
 * sqpqp.h -- C ABI of the B200-native QP-subproblem engine for SqpSolver.jl.
 *
 * This is the drop-in boundary: the entry points below are what SqpSolver.jl's
 * Julia host would bind with `ccall` (see INTEGRATION.md and
 * sqpsolver.jl_b200/julia/SqpQpB200.jl) in place of the JuMP->MOI->Ipopt path the
 * reference takes in src/algorithms/subproblem_JuMP.jl.  Each function cites the
 * reference code it replaces.
 *
 * Conventions
 *   - return 0 = OK, < 0 = error (see SQPQP_E_*); solver outcome is DATA
 *     (`moi_status`), never an error code.
 *   - no C++ exceptions, no callbacks cross the boundary.
 *   - all pointers are HOST pointers owned by the caller unless the function name
 *     ends in `_device`; they are only read/written during the call.
 *   - all device memory, pinned staging buffers and the CUDA stream are owned by
 *     the handle.  A handle is not thread-safe; different handles are independent.
 *   - per-instance arrays of a batch are instance-major: a[b*len + i].
 *     A single NLP is a batch of 1.
 *   - indices in COO inputs are 1-based int64 exactly as MOI hands them to
 *     SqpSolver (src/MOI_wrapper.jl:930-945, 1010-1025).
 *   - multipliers come back in the reference's storage convention
 *     (subproblem_JuMP.jl:514-563): lambda in MOI sign (grad = J' lambda + r),
 *     mult_x_L = max(r,0) >= 0, mult_x_U = min(r,0) <= 0.
 */
#ifndef SQPQP_H
#define SQPQP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sqpqp_handle_s* sqpqp_handle;

/* error codes */
#define SQPQP_OK 0
#define SQPQP_E_BADARG (-1)
#define SQPQP_E_CUDA (-2)
#define SQPQP_E_NOMEM (-3)
#define SQPQP_E_STATE (-4) /* call order violated (e.g. solve before setup/update) */

/* moi_status values == Integer(MOI.TerminationStatusCode) in MathOptInterface v1 */
#define SQPQP_MOI_OPTIMIZE_NOT_CALLED 0
#define SQPQP_MOI_OPTIMAL 1
#define SQPQP_MOI_INFEASIBLE 2
#define SQPQP_MOI_LOCALLY_SOLVED 4
#define SQPQP_MOI_LOCALLY_INFEASIBLE 5
#define SQPQP_MOI_ALMOST_LOCALLY_SOLVED 10
#define SQPQP_MOI_ITERATION_LIMIT 11
#define SQPQP_MOI_NUMERICAL_ERROR 20
/* Status semantics (what the SQP driver branches on, sqp_trust_region.jl:144-178):
 *   LOCALLY_SOLVED         primal and dual residual <= ipm_eps (1e-9, relative to max(1,|x|,|Ax|) resp. the gradient
 *                          terms) and complementarity <= ipm_eps; or the verified active-set (KKT) refinement of ADMM.
 *   ALMOST_LOCALLY_SOLVED  Ipopt's "solved to acceptable level": both residuals and the complementarity <= 1e-6 on
 *                          the same relative scales (either path), reached at the iteration cap or on the floor of an
 *                          acceptable point.  Never granted at a looser level.
 *   LOCALLY_INFEASIBLE     certificate on the multiplier direction or a stalled primal residual (zero-filled outputs).
 *   ITERATION_LIMIT        the iterate is returned (never uninitialised memory) but is NOT in the driver's ok set.
 *   NUMERICAL_ERROR        no factorisation / no ADMM progress; outputs zero-filled. */

/* phases of sqpqp_solve_tr */
#define SQPQP_PHASE_QP 0  /* sub_optimize!      subproblem_JuMP.jl:127-183 */
#define SQPQP_PHASE_FR 1  /* sub_optimize_FR!   subproblem_JuMP.jl:352-393 */
#define SQPQP_PHASE_SOC 2 /* sub_optimize_soc!  sqp_trust_region.jl:341-360 (E_override = g(x+p) - J p) */
#define SQPQP_PHASE_LP 3  /* sub_optimize_lp    subproblem_JuMP.jl:185-244 (start-point projection) */

/* Per-instance solve statistics (all counters are exact, measured on device). */
typedef struct sqpqp_info {
    int32_t moi_status;
    int32_t admm_iters;
    int32_t cg_iters;        /* PCG iterations inside ADMM x-updates */
    int32_t polish_tries;
    int32_t polish_cg_iters; /* PCG iterations inside polish solves */
    int32_t polished;        /* 1: returned point is the verified active-set (KKT) refinement */
    int32_t rho_updates;
    int32_t checks;          /* residual checks (3 SpMV each) */
    int32_t ipm_iters;       /* interior-point iterations (0 if the ADMM path solved it) */
    int32_t chol_factorizations; /* sparse Cholesky factorisations of the condensed Newton matrix */
    double rho;              /* final ADMM step size (scaled space) */
    double rho_box_floor;    /* 1.5*|lambda_min(P_scaled)| -- nonconvexity guard, 0 if P >= 0 */
    double res_prim;         /* unscaled inf-norm primal residual of the returned point */
    double res_dual;         /* unscaled inf-norm dual residual of the returned point */
    double objective;        /* 1/2 p'Pp + q'p at the returned point */
} sqpqp_info;

/* Solver options (defaults via sqpqp_default_options). */
typedef struct sqpqp_options {
    double rho0, sigma, alpha;
    double eps_abs, eps_rel, eps_inf;
    double rho_eq_mult, rho_min, rho_max, adapt_tol;
    double cg_rel0;          /* PCG stops at cg_rel0 * |residual of the warm start| */
    double rb_full_mult;     /* rho_box floor = rb_full_mult * max(0,-lambda_min(P_scaled)) */
    double polish_trigger, polish_rho, polish_tol, feas_tol, dual_tol;
    int32_t max_iter, check_every, ruiz_iters, cg_max, eig_iters, polish_outer, polish_cg_max;
    int32_t warm_start;      /* 1: start ADMM from the previous solve's (p, y) of the same instance */
    int32_t team;            /* 0 auto, 1 one CTA per instance, 2 whole grid per instance (cooperative) */
    int32_t threads;         /* CTA size (multiple of 32, <= 512); 0 = auto */
    int32_t method;          /* 0 auto: interior point (Cholesky) with ADMM fallback; 1 ADMM only; 2 interior point only */
    int32_t ipm_max_iter;
    int32_t fallback_max_iter; /* ADMM iteration cap when it runs as the fallback of a failed interior-point solve */
    double ipm_eps, ipm_delta0, ipm_delta_min, ipm_rho0, ipm_tau, ipm_mu0, ipm_mu_min, ipm_kappa_eps;
    int32_t ipm_refine;      /* reserved (iterative refinement of the Newton solve was measured useless and removed) */
    int32_t verbose;         /* 1: device printf of the interior-point iterations (debugging) */
    int32_t occupancy;       /* CTA kernel variant: 0 auto (2 for large batches, 1 for small), 1 = one 512-thread CTA per SM with
                                128 regs/thread and resident work vectors, 2 = two CTAs per SM (384 threads / 80 regs by default, 512 / 64 on request),
                                >= 3 = four 256-thread CTAs per SM (64 regs) */
    int32_t smem_kb;         /* shared-memory budget per CTA for resident work vectors; -1 = auto, 0 = none (the dense tail of
                                the factor, its inverse diagonal and the solve scratch are always resident) */
    double ipm_ic_growth;    /* inertia correction: factor on the diagonal shift after a failed factorisation (default 4) */
    double ipm_ic_decay;     /* ... and divisor of the shift per interior-point iteration once it has succeeded (default 3) */
} sqpqp_options;

/* ---- lifecycle ------------------------------------------------------------------ */
/* One handle = one CUDA device + one stream + owned device/pinned buffers. */
int sqpqp_create(sqpqp_handle* h, int device_ordinal);
int sqpqp_destroy(sqpqp_handle h);
const char* sqpqp_last_error(sqpqp_handle h);
void sqpqp_default_options(sqpqp_options* o);
int sqpqp_set_options(sqpqp_handle h, const sqpqp_options* o);
/* CUDA stream of the handle as a void* (cudaStream_t), for callers that time or
 * enqueue work on the same stream. */
void* sqpqp_stream(sqpqp_handle h);
/* Page-locks a caller-owned host buffer once (cudaHostRegister) and remembers its range.  Every later call of the NLP lane
 * that is handed a pointer inside a registered range copies directly between that buffer and the device instead of
 * through the library's pinned staging area (measured on the B200 box: 55 GB/s instead of ~21 GB/s).  Calls stay blocking:
 * the buffer is free for the caller on return.  The reference's host keeps dE, h_val, df, E, x, lambda, mult_x_* as
 * persistent vectors of the SQP object (sqp.jl:16-59) -- register those once after they are allocated.  Unregister before
 * freeing the memory; sqpqp_destroy releases whatever is still registered.  A buffer the caller page-locked itself is
 * accepted (only its range is recorded). */
int sqpqp_host_register(sqpqp_handle h, void* ptr, int64_t bytes);
int sqpqp_host_unregister(sqpqp_handle h, void* ptr);

/* ---- NLP lane (fast lane B2: QpDevice <: AbstractSubOptimizer) --------------------- */
/* Replaces the SqpTR constructor's pattern build (sqp_trust_region.jl:41-57:
 * sparse(j_row,j_col,ones,m,n), sparse(h_row,h_col,ones,n,n)) and create_model!
 * (subproblem_JuMP.jl:36-125: slack columns for rows > m_lin, 2 per two-sided row).
 * Builds on device, once: CSR(J | slack cols), CSR of its transpose (== Julia's CSC
 * of J), the symmetric-full CSR(H) of sqp.jl:92-103, and the ordered-duplicate
 * scatter permutations.  `batch` instances share the pattern and x_L/x_U/g_L/g_U
 * finiteness; bounds are given per instance when bounds_per_instance != 0
 * (instance-major), else once.  +-Inf allowed in bounds. */
int sqpqp_setup_nlp(sqpqp_handle h, int32_t batch, int32_t n, int32_t m, int32_t m_lin,
                    int64_t nnz_j, const int64_t* j_row, const int64_t* j_col,
                    int64_t nnz_h, const int64_t* h_row, const int64_t* h_col,
                    const double* x_L, const double* x_U, const double* g_L, const double* g_U,
                    int32_t bounds_per_instance);

/* Replaces eval_functions!/eval_Jacobian! scatter (sqp.jl:86-117) and the QpData
 * refresh (sqp.jl:66-79): dE[batch][nnz_j], h_val[batch][nnz_h] (may be NULL: no
 * Hessian), df[batch][n], E[batch][m], lambda is NOT needed (already folded in h_val).
 * Host arrays are copied to pinned staging and uploaded asynchronously; duplicates
 * are summed in ascending COO order starting from 0.0 (bit-exact with the reference). */
int sqpqp_update_nlp(sqpqp_handle h, const double* dE, const double* h_val, const double* df, const double* E);
/* Same, with DEVICE pointers (no staging, no copy): for a device-side evaluator. */
int sqpqp_update_nlp_device(sqpqp_handle h, const double* dE, const double* h_val, const double* df, const double* E);

/* Replaces sub_optimize! / sub_optimize_FR! / sub_optimize_soc! / sub_optimize_lp +
 * set_trust_region! + modify_constraints! + collect_solution!
 * (subproblem_JuMP.jl:127-183, 352-393, 185-244, 432-448, 465-512, 514-563).
 *   x_k[batch][n], delta[batch]; E_override[batch][m] only for SQPQP_PHASE_SOC else NULL;
 *   active[batch] (may be NULL = all): instances with active[b]==0 are skipped and
 *   their outputs left untouched.
 * Outputs (any may be NULL): p[batch][n] (for PHASE_LP: the projected x),
 *   lambda[batch][m], mult_x_L/U[batch][n], slack[batch][S], moi_status[batch],
 *   info[batch].  Blocking. */
int sqpqp_solve_tr(sqpqp_handle h, int32_t phase, const double* x_k, const double* delta,
                   const double* E_override, const int32_t* active,
                   double* p, double* lambda, double* mult_x_L, double* mult_x_U, double* slack,
                   int32_t* moi_status, sqpqp_info* info);
/* Device-pointer variant for a device-side evaluator / graph of solves: all arguments are
 * DEVICE pointers, nothing is copied, the call does not block.  Results stay in the handle's
 * device buffers (sqpqp_device_outputs); sqpqp_sync waits for the stream; sqpqp_fetch_info
 * copies only the per-instance info records back. */
int sqpqp_solve_tr_device(sqpqp_handle h, int32_t phase, const double* x_k, const double* delta,
                          const double* E_override, const int32_t* active);
/* One SQP round of a BATCH whose instances are in different phases (compute_step!, sqp_trust_region.jl:370-380, decides per
 * instance between sub_optimize! and sub_optimize_FR!): instances with active_qp[b] != 0 solve the QP subproblem, instances with
 * active_fr[b] != 0 the feasibility-restoration LP, all others are skipped (outputs untouched).  The two masks must be disjoint
 * (SQPQP_E_BADARG otherwise).  Equivalent to two sqpqp_solve_tr calls, but x_k / delta go up once, the results come back once,
 * and the two launches run side by side on two streams instead of one after the other (a round has few instances in
 * restoration; their launch alone leaves the GPU idle).  Blocking; outputs as for sqpqp_solve_tr. */
int sqpqp_solve_tr_mixed(sqpqp_handle h, const double* x_k, const double* delta, const int32_t* active_qp,
                         const int32_t* active_fr, double* p, double* lambda, double* mult_x_L, double* mult_x_U,
                         double* slack, int32_t* moi_status, sqpqp_info* info);
/* ... with DEVICE pointers (masks included; the caller guarantees that they are disjoint); does not block. */
int sqpqp_solve_tr_mixed_device(sqpqp_handle h, const double* x_k, const double* delta, const int32_t* active_qp,
                                const int32_t* active_fr);
int sqpqp_sync(sqpqp_handle h);
int sqpqp_device_outputs(sqpqp_handle h, double** p, double** lambda, double** mult_x_L, double** mult_x_U,
                         sqpqp_info** info);
int sqpqp_fetch_info(sqpqp_handle h, sqpqp_info* info);
/* Size of the shared symbolic Cholesky factor of the condensed Newton matrix (0 if unavailable):
 * nnz(L), level-scheduled (sparse) elimination-tree levels, flops of the sparse part per numeric
 * factorisation. */
int sqpqp_chol_stats(sqpqp_handle h, int64_t* nnzL, int64_t* nlev, int64_t* flops);
/* Layout of that factor: columns of the dense tail (top of the elimination tree, factorised as a
 * packed dense matrix in shared memory; 0 = none) and levels of the whole elimination tree. */
int sqpqp_chol_layout(sqpqp_handle h, int64_t* tail_cols, int64_t* tree_levels);
/* Development aid: cycles per solve segment summed over CTAs since the last call (32 counters; all
 * zero unless the library was built with -DSQPQP_PROF). */
int sqpqp_prof_read(sqpqp_handle h, uint64_t* out32);
/* Development aid: copy a per-instance work array of instance b to the host (kind 0 N-vector slot idx, 1 M-vector
 * slot idx, 2 factor values, 3 solve scratch, 4 inverse diagonal of the factor, 5 weighted Jacobian values). */
int sqpqp_debug_set(sqpqp_handle h, int32_t what, int32_t value); /* what 0: resident SpMV CTAs per SM; 1: cap of the dense
                                                                     tail of the factor in columns (-1 = auto), takes
                                                                     effect at the next setup */
int sqpqp_debug_read(sqpqp_handle h, int32_t kind, int32_t idx, int32_t b, double* out, int64_t count);
/* Launch order of the batched solves that follow: CTA slot k of a launch runs instance order[k] (host array, a permutation
 * of 0..batch-1; NULL restores index order).  The hardware starts CTAs in slot order, so a caller that knows which
 * instances are slow (e.g. from its own history) can start them first.  Results are independent of the order. */
int sqpqp_set_launch_order(sqpqp_handle h, const int32_t* order);
/* Development aid: raw interior-point loop states (and the per-instance flags behind them) saved by a launch that ran
 * with an iteration quota (sqpqp_debug_set what = 8 / 9). */
int sqpqp_debug_read_state(sqpqp_handle h, void* out, int64_t bytes);
/* Number of slack columns S (order: for each row i > m_lin: u_i, then v_i if two-sided). */
int sqpqp_num_slacks(sqpqp_handle h, int32_t* S);

/* Merit / model arithmetic of the ratio test on device:
 *   norm_violations (common.jl:54-77, p=1), compute_phi (sqp.jl:170-183),
 *   compute_qmodel (sqp_trust_region.jl:487-508).
 * Inputs per instance: x[n], p[n], E_trial[m] = g(x+p) and f_trial = f(x+p) from the
 * host callbacks, mu, fr (feasibility-restoration flag).  Uses the E, df, J, H of the
 * last update_nlp.  Outputs per instance (any may be NULL):
 *   viol0 = |viol(E,x)|_1, viol_trial = |viol(E_trial,x+p)|_1,
 *   phi_trial = f_trial + mu*viol_trial (or viol_trial if fr),
 *   q0 = mu*viol0, qk = df'p + 1/2 p'Hp + mu*|viol(E+Jp, x+p)|_1. */
int sqpqp_merit(sqpqp_handle h, const double* x, const double* p, const double* E_trial, const double* f_trial,
                const double* mu, const int32_t* fr,
                double* viol0, double* viol_trial, double* phi_trial, double* q0, double* qk);
/* Device-side evaluator of the polar ACOPF NLP for the batched workload (csrc/acopf.cuh): replaces the MOI NLPEvaluator
 * callbacks of eval_functions! (sqp.jl:86-104) and the upload of their results.  The network must be the one whose
 * structure was given to sqpqp_setup_nlp (formulation and COO order: sqpsolver.jl_b200/nlp/acopf.py = PowerModels
 * ACPPowerModel + build_opf of the reference's test/opf.jl:5-9).  oa/oc/os are the [nl][4] Ohm-row coefficients, the
 * balance rows are given as CSR over their Jacobian COO entries (kind 0 constant coefficient, 1 P shunt, 2 Q shunt).
 * sqpqp_acopf_eval_update evaluates f, grad f, g, the Jacobian and Lagrangian-Hessian (sigma = 1, mu = lambda) COO
 * values of the instances with mask != 0 (all if null) at x, scatters them, and returns f[batch], E[batch][m],
 * df[batch][n]; instances not in the mask keep their previous values. */
int sqpqp_acopf_setup(sqpqp_handle h, int32_t nb, int32_t ng, int32_t nl, int32_t ref_bus, const int32_t* f_bus,
                      const int32_t* t_bus, const int32_t* gen_bus, const double* oa, const double* oc, const double* os,
                      const double* cost2, const double* cost1, const double* cost0, const double* gs, const double* bs,
                      int32_t nbal, const int32_t* bal_ptr, const int32_t* bal_col, const int32_t* bal_kind,
                      const double* bal_const, int32_t nsh, const int32_t* sh_bus);
int sqpqp_acopf_eval_update(sqpqp_handle h, const double* x, const double* lambda, const int32_t* mask, double* f, double* E,
                            double* df);
/* f and g at a TRIAL point on the device (compute_phi with alpha > 0, sqp.jl:170-183, as called by do_step!,
 * sqp_trust_region.jl:515-530): function values only, kept in device-side trial buffers; a following
 * sqpqp_merit(..., E_trial = NULL, f_trial = NULL, ...) reads them in place, so neither array crosses the link.
 * mask (nullable): instances with 0 keep a copy of their current E and f.  f / E may be NULL. */
int sqpqp_acopf_eval_trial(sqpqp_handle h, const double* x_trial, const int32_t* mask, double* f, double* E);
/* Line-search primitives on the current device matrices (the reference's line-search driver, sqp_line_search.jl, is not
 * compiled by the reference -- sqp.jl:226 -- these are the device quantities its merit maths needs: compute_mu_rule2!
 * :280-291, compute_alpha :303-334, compute_phi sqp.jl:170-183, compute_derivative sqp.jl:190-213 + merit.jl:13-17,
 * norm_complementarity common.jl:30-47).  Inputs per instance: x[n], p[n], alpha, E_trial[m] = g(x + alpha p),
 * mu_rows[m], lambda[m].  out8 is [8][batch]:
 *   0 df'p   1 p'Hp   2 |viol(E,x)|_1   3 |viol(E,x)|_inf   4 sum_i mu_i viol_i(E) + |mu|_inf sum_j viol_j(x)
 *   5 the same at (E_trial, x + alpha p)   6 |viol(E_trial, x + alpha p)|_1   7 norm_complementarity(E, lambda), p = Inf. */
int sqpqp_linesearch_terms(sqpqp_handle h, const double* x, const double* p, const double* alpha, const double* E_trial,
                           const double* mu_rows, const double* lambda, double* out8);
/* KT_residuals (common.jl:14-23) as coded, per instance. */
int sqpqp_kt_residuals(sqpqp_handle h, const double* lambda, const double* mult_x_U, const double* mult_x_L,
                       double* kt);

/* out[batch][m] = J p per instance (`Jacobian * p`, sqp_trust_region.jl:343, 492). */
int sqpqp_jac_times(sqpqp_handle h, const double* p, double* out);
/* Batched CSR sparse matrix-vector products on the current device matrices, one shared pattern
 * (csrc/spmv.cuh, CSR-stream): which = 0: y[batch][m] = J x (sqp_trust_region.jl:343,492),
 * 1: y[batch][n] = J' x (common.jl:17), 2: y[batch][n] = H x (sqp_trust_region.jl:490).
 * sqpqp_spmv takes host buffers and blocks; sqpqp_spmv_device takes device pointers, does not
 * block, and is timed by sqpqp_last_solve_ms. */
int sqpqp_spmv(sqpqp_handle h, int32_t which, const double* x, double* y);
int sqpqp_spmv_device(sqpqp_handle h, int32_t which, const double* x_dev, double* y_dev);

/* Read back the device matrices of instance b (parity tests): CSR of J (m x n, slack
 * columns excluded), CSR of J' (n x m) and symmetric CSR of H.  Pass NULL to skip.
 * which: 0 = J, 1 = J transposed, 2 = H.  nnz query: pass row_ptr=NULL. */
int sqpqp_get_csr(sqpqp_handle h, int32_t which, int32_t b, int64_t* nnz, int32_t* row_ptr, int32_t* col_idx,
                  double* values);

/* ---- generic QP lane (boundary B1: what an MOI.AbstractOptimizer shim sees) -------- */
/*   min 1/2 x'Px + q'x   s.t.  rl <= A x <= ru,  cl <= x <= cu
 * P is given as MOI ScalarQuadraticTerm triplets (one triangle; an off-diagonal
 * term c means P_ij = P_ji = c, a diagonal term c means P_ii = c; duplicates add),
 * A as 1-based triplets.  Equivalent to the NLP lane with m_lin = 0, E = 0, x_k = 0,
 * delta = +Inf. */
int sqpqp_qp_setup(sqpqp_handle h, int32_t nv, int32_t nc, int64_t nnz_p, const int64_t* p_row, const int64_t* p_col,
                   int64_t nnz_a, const int64_t* a_row, const int64_t* a_col);
int sqpqp_qp_solve(sqpqp_handle h, const double* p_val, const double* q, const double* a_val,
                   const double* rl, const double* ru, const double* cl, const double* cu,
                   double* x, double* row_dual, double* col_dual, int32_t* moi_status, sqpqp_info* info);

/* ---- introspection ----------------------------------------------------------------- */
/* Kernel launches issued by this handle since creation (bench.py's gpu_launches). */
int64_t sqpqp_launch_count(sqpqp_handle h);
/* Device time (ms) of the last sqpqp_solve_tr's solve kernel, from CUDA events on
 * the handle's stream. */
double sqpqp_last_solve_ms(sqpqp_handle h);
/* Sum of the CUDA-event durations of every solve launch of this handle so far (one SQP step may launch the QP phase and
 * the restoration phase; sqpqp_last_solve_ms sees only the last of them). */
double sqpqp_solve_ms_total(sqpqp_handle h);
/* Name and launch shape of the interior-point kernel the last solve launched (e.g. "k_solve_cta<384,2,1>",
 * "k_solve_ilv<4,512,1>", "k_solve_grid"): the launch rule lives in the library, reports read it from here. */
const char* sqpqp_last_solve_kernel(sqpqp_handle h);

#ifdef __cplusplus
}
#endif
#endif /* SQPQP_H */
