// sqpqp.cu -- C ABI (include/sqpqp.h) of the B200-native QP-subproblem engine.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
//
// Host side of the boundary: handle/stream/arena management, pinned staging for the
// per-iteration value upload, the one-time device pattern build (K1) and kernel launches.
// There is deliberately NO CPU implementation of any numerical step here: every entry point
// that computes launches a kernel on the handle's stream.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <limits>

#include "common.cuh"
#include "pattern.cuh"
#include "admm.cuh"
#include "ipm.cuh"
#include "ilv.cuh"
#include "merit.cuh"
#include "spmv.cuh"
#include "acopf.cuh"
#include "symbolic.hpp"

struct Pending {
    const void* pin;
    void* user;
    size_t bytes;
    cudaEvent_t ev;  // recorded behind the D2H copy of a large item (null: wait for the stream)
};

struct sqpqp_handle_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // pair in use (one of tev[], or the SpMV pair)
    cudaEvent_t tev[4][2] = {};                 // ring of event pairs of the solve launches
    cudaEvent_t sev[2] = {};                    // pair of the SpMV timing
    cudaStream_t stream2 = nullptr;             // side stream (highest priority) of the mixed-phase launch: the restoration-phase kernel
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;  // runs next to the QP-phase kernel of the same round (launch_solve_mixed)
    int tp_head = 0, tp_n = 0;
    std::string err;
    sqpqp_options opts;
    bool setup_done = false, updated = false;
    Prob P;
    int64_t nnzJ_coo = 0, nnzH_coo = 0;
    std::vector<void*> allocs;
    // staging (pinned host + device mirror), bump-allocated per API call
    char* pin = nullptr;
    char* dstage = nullptr;
    size_t stage_cap = 0, stage_off = 0;
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> evpool;  // per-item completion events of the staged downloads
    size_t ev_used = 0;
    // scatter
    ScatterJob jobJ{}, jobT{}, jobH{};
    double *d_dE = nullptr, *d_hval = nullptr, *d_df = nullptr, *d_E = nullptr;  // owned copies
    double *d_gL = nullptr, *d_gU = nullptr, *d_xL = nullptr, *d_xU = nullptr;
    double *d_xk = nullptr, *d_delta = nullptr, *d_Eov = nullptr;
    int* d_active = nullptr;
    int64_t launches = 0;
    cudaError_t async_err = cudaSuccess;  // first failure of a staged copy (upload / download); reported by finish() / the caller
    int tail_override = -1;               // development knob (sqpqp_debug_set what = 1): cap of the dense tail in columns
    int handoff = -1;                     // what = 8: iteration quota of the throughput launch before an instance is handed to the
                                          // resident launch; -1 auto (40 when num_sms < batch <= 2 num_sms and the ring fits), 0 off
    int fuse_fwd = 1;                     // what = 7: 0 keeps the forward sweep out of the factor program (slot lists); at setup
    // what = 9: what follows the quota of the throughput launch.  0: the resident ring launch continues the stragglers
    // (shards of one wave); 1: a second THROUGHPUT launch continues every unfinished instance LONGEST-FIRST, in the order
    // k_rank_unfinished predicts from the saved interior-point state (multi-wave shards: the tail of a launch is its
    // stragglers, and which instances they are shows in the first iterations); 2: nothing (development: state read-back);
    // 3: second throughput launch in index order (control of 1)
    int handoff_mode = 0;
    int* d_order = nullptr;               // [batch] launch order of the second stage / of sqpqp_set_launch_order
    bool order_user = false;              // d_order was set by the caller and applies to the next launches
    int ring_enable = 1;                  // sqpqp_debug_set what = 5: 0 keeps the slot lists (no ring programs are built); at setup
    bool last_ring = false;
    // caller buffers page-locked with sqpqp_host_register: copied to / from the device directly, without the pinned staging hop
    std::vector<std::pair<const char*, size_t>> host_pinned;
    std::vector<const char*> host_owned_reg;  // the subset this handle page-locked itself (and must unlock)
    int ring_mode = 0;                    // what = 6: launches that stream the ring: 0 auto, 1 never, 2 whenever the programs exist
    double last_ms = 0.0;
    bool timing_pending = false;
    bool timing_is_solve = false;         // the pending event pair brackets a solve launch (else an SpMV)
    double solve_ms_total = 0.0;          // CUDA-event time of every solve launch so far (sqpqp_solve_ms_total)
    int num_sms = 148, coop_blocks = 0, max_dyn_smem = 0;
    int chol_nnzL = 0, chol_nlev = 0, chol_tail = 0, chol_nlev_total = 0;
    int cta4_smem = 48 * 1024;  // dynamic shared memory of one CTA when four share an SM
    int cta2_smem = 100 * 1024; // ... when two share an SM (the default for large batches)
    int64_t chol_flops = 0;
    std::vector<double> avg_row;  // avg row length of J(normal), J(ext), T, H
    SpmvPlan planJ{}, planT{}, planH{};
    int spmv_ctas_per_sm = 8;
    AcopfDev acopf{};       // device-side ACOPF evaluator (acopf.cuh); nb == 0: not set up
    double* d_f = nullptr;  // [batch] objective values of the evaluator
    double *d_Etrial = nullptr, *d_ftrial = nullptr;  // trial-point values of sqpqp_acopf_eval_trial
    bool trial_valid = false;  // CSR-stream row blocks of J (normal phase), J' and H
    // interleaved batch path (ilv.cuh): G instances per CTA.  ilv_G: 0 = auto, 1 = off, 2 / 4 / 8; chosen at setup
    int ilv_G = 0, ilv_threads = 0, ilv_occ = 0;
    int G = 1;                       // in effect for the current problem
    IlvDev ilv{}, ilv_fr{};
    bool has_ilv = false, has_ilv_fr = false;
    size_t ilv_dyn = 0, ilv_fr_dyn = 0;
    int ilv_nt = 512, ilv_minb = 1;
    std::string last_kernel;         // name of the interior-point kernel of the last solve launch (sqpqp_last_solve_kernel)
    // generic-lane bookkeeping
    bool generic = false;
};

static int fail(sqpqp_handle h, int code, const char* msg) {
    if (h) h->err = msg;
    return code;
}
static int fail_cuda(sqpqp_handle h, cudaError_t e, const char* what, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %s at line %d: %s", cudaGetErrorString(e), line, what);
    if (h) h->err = buf;
    return (e == cudaErrorMemoryAllocation) ? SQPQP_E_NOMEM : SQPQP_E_CUDA;
}

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() { cudaSetDevice(prev); }
};

template <class T>
static int dalloc(sqpqp_handle h, T** p, size_t count) {
    if (count == 0) count = 1;
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T));
    if (e != cudaSuccess) return fail_cuda(h, e, "cudaMalloc", __LINE__);
    e = cudaMemsetAsync(q, 0, count * sizeof(T), h->stream);
    if (e != cudaSuccess) return fail_cuda(h, e, "cudaMemsetAsync", __LINE__);
    h->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}
#define DALLOC(ptr, count)                                  \
    do {                                                    \
        int rc__ = dalloc(h, &(ptr), (size_t)(count));      \
        if (rc__) return rc__;                              \
    } while (0)

static int ensure_stage(sqpqp_handle h, size_t bytes) {
    bytes += 4096;
    if (bytes > h->stage_cap) {
        CUDA_OK(cudaStreamSynchronize(h->stream));
        if (h->pin) cudaFreeHost(h->pin);
        if (h->dstage) cudaFree(h->dstage);
        h->pin = nullptr;
        h->dstage = nullptr;
        size_t cap = bytes + bytes / 4;
        CUDA_OK(cudaMallocHost((void**)&h->pin, cap));
        CUDA_OK(cudaMalloc((void**)&h->dstage, cap));
        h->stage_cap = cap;
    }
    h->stage_off = 0;
    h->pending.clear();
    h->ev_used = 0;
    h->async_err = cudaSuccess;
    return 0;
}
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// host -> pinned -> device (async).  Returns device pointer inside the staging mirror.
// Host-side copy between caller memory and the pinned staging area.  Above a few MB one thread cannot feed PCIe
// (~8 GB/s single-threaded memcpy): split the copy over a few threads.  Measured on the B200 box (tools/
// gpu_e2e_breakdown.py): 134 MB of update_nlp inputs take 6.4 ms = 21 GB/s whatever the piece size and thread count
// (2..14 threads, 16 MB..whole array) -- the link, not the staging copy, is the floor.
static void par_memcpy(void* dst, const void* src, size_t bytes) {
    const size_t kMin = (size_t)2 << 20, kMaxT = 8;
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = bytes / kMin;
    if (nt > kMaxT) nt = kMaxT;
    if (hw && nt > hw) nt = hw;
    if (nt < 2) { memcpy(dst, src, bytes); return; }
    std::vector<std::thread> th;
    const size_t chunk = ((bytes / nt) + 4095) & ~(size_t)4095;
    for (size_t t = 1; t < nt; ++t) {
        size_t off = t * chunk;
        if (off >= bytes) break;
        size_t len = bytes - off < chunk ? bytes - off : chunk;
        th.emplace_back([=] { memcpy((char*)dst + off, (const char*)src + off, len); });
    }
    memcpy(dst, src, chunk < bytes ? chunk : bytes);
    for (auto& x : th) x.join();
}

// Staged host -> device copy, pipelined: the caller's buffer is copied into the pinned area in pieces and every piece
// starts its H2D transfer as soon as it is staged, so the transfer overlaps the host copy of the next piece.
static bool is_registered(sqpqp_handle h, const void* p, size_t bytes) {
    const char* c = (const char*)p;
    for (auto& r : h->host_pinned)
        if (c >= r.first && c + bytes <= r.first + r.second) return true;
    return false;
}
static const size_t kStagePiece = (size_t)32 << 20;
// one stage of the index-program ring (chol.cuh): header + 512 slots + 4 pair words per slot
static const int kTimingRing = 4;
static const int kRingStageBytes = 32 + 512 * 8 + 4 * 512 * 4;
static cudaError_t stage_h2d(sqpqp_handle h, void* ddst, size_t pin_off, const void* src, size_t bytes) {
    // a registered (page-locked) caller buffer goes over the link as it is: the staging memcpy tops out near 21 GB/s on the
    // B200 box, the link itself gives 55 GB/s (tools/gpu_pcie.py)
    if (is_registered(h, src, bytes)) return cudaMemcpyAsync(ddst, src, bytes, cudaMemcpyHostToDevice, h->stream);
    for (size_t o = 0; o < bytes; o += kStagePiece) {
        const size_t len = bytes - o < kStagePiece ? bytes - o : kStagePiece;
        par_memcpy(h->pin + pin_off + o, (const char*)src + o, len);
        cudaError_t e = cudaMemcpyAsync((char*)ddst + o, h->pin + pin_off + o, len, cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template <class T>
static const T* upload(sqpqp_handle h, const T* src, size_t count) {
    if (!src || count == 0) return nullptr;
    size_t bytes = count * sizeof(T);
    size_t off = h->stage_off;
    if (off + bytes > h->stage_cap) {  // the caller sized the staging area too small: never write past it
        if (h->async_err == cudaSuccess) h->async_err = cudaErrorInvalidValue;
        return nullptr;
    }
    h->stage_off = align256(off + bytes);
    cudaError_t e = stage_h2d(h, h->dstage + off, off, src, bytes);
    if (e != cudaSuccess && h->async_err == cudaSuccess) h->async_err = e;
    return (const T*)(h->dstage + off);
}
// device -> pinned (async) and remember the final host copy; large items get their own completion event so that
// finish() can copy item k to the caller while item k+1 is still crossing PCIe
template <class T>
static void download(sqpqp_handle h, const T* dsrc, T* user, size_t count) {
    if (!user || count == 0) return;
    size_t bytes = count * sizeof(T);
    size_t off = h->stage_off;
    if (off + bytes > h->stage_cap) {
        if (h->async_err == cudaSuccess) h->async_err = cudaErrorInvalidValue;
        return;
    }
    if (is_registered(h, user, bytes)) {  // straight into the caller's page-locked buffer; finish() waits for the stream
        cudaError_t ce = cudaMemcpyAsync(user, dsrc, bytes, cudaMemcpyDeviceToHost, h->stream);
        if (ce != cudaSuccess && h->async_err == cudaSuccess) h->async_err = ce;
        return;
    }
    h->stage_off = align256(off + bytes);
    cudaError_t ce = cudaMemcpyAsync(h->pin + off, dsrc, bytes, cudaMemcpyDeviceToHost, h->stream);
    if (ce != cudaSuccess && h->async_err == cudaSuccess) h->async_err = ce;
    cudaEvent_t ev = nullptr;
    if (bytes >= ((size_t)1 << 20)) {
        if (h->ev_used == h->evpool.size()) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess) h->evpool.push_back(e);
        }
        if (h->ev_used < h->evpool.size()) {
            ev = h->evpool[h->ev_used++];
            cudaEventRecord(ev, h->stream);
        }
    }
    h->pending.push_back(Pending{h->pin + off, user, bytes, ev});
}
static int finish(sqpqp_handle h) {
    if (h->async_err != cudaSuccess) {  // a staged copy of this call failed: report it instead of handing back stale data
        cudaError_t e = h->async_err;
        h->async_err = cudaSuccess;
        cudaStreamSynchronize(h->stream);
        h->pending.clear();
        h->ev_used = 0;
        return fail_cuda(h, e, "staged host<->device copy", __LINE__);
    }
    for (auto& p : h->pending)
        if (p.ev) {
            CUDA_OK(cudaEventSynchronize(p.ev));
            par_memcpy(p.user, p.pin, p.bytes);
        }
    CUDA_OK(cudaStreamSynchronize(h->stream));
    for (auto& p : h->pending)
        if (!p.ev) par_memcpy(p.user, p.pin, p.bytes);
    h->pending.clear();
    h->ev_used = 0;
    CUDA_OK(cudaGetLastError());
    return 0;
}

static int grid_for(int64_t work, int threads) {
    int64_t g = (work + threads - 1) / threads;
    if (g < 1) g = 1;
    if (g > 148 * 16) g = 148 * 16;
    return (int)g;
}
static int lg_lanes(double avg) {
    int lg = 0;
    while ((1 << lg) < avg && lg < 5) ++lg;
    return lg;
}

// ---- K1 driver ---------------------------------------------------------------------------
struct Pattern {
    int nrows = 0, nslots = 0, Lv = 0;
    int *row_ptr = nullptr, *col_idx = nullptr, *seg_ptr = nullptr, *seg_src = nullptr;
};
static int build_pattern(sqpqp_handle h, int L, const int* erow, const int* ecol, const int* esrc, int nrows, Pattern* out) {
    int *cnt, *rstart, *cursor, *tcol, *tent, *scol, *sent, *head, *slotof;
    DALLOC(cnt, nrows + 1);
    DALLOC(rstart, nrows + 1);
    DALLOC(cursor, nrows + 1);
    DALLOC(tcol, L);
    DALLOC(tent, L);
    DALLOC(scol, L);
    DALLOC(sent, L);
    DALLOC(head, L + 1);
    DALLOC(slotof, L + 2);
    DALLOC(out->row_ptr, nrows + 1);
    DALLOC(out->col_idx, L);
    DALLOC(out->seg_ptr, L + 1);
    DALLOC(out->seg_src, L);
    const int TB = 256;
    if (L > 0) {
        k_count_rows<<<grid_for(L, TB), TB, 0, h->stream>>>(L, erow, cnt);
        k_exscan<<<1, 1024, 0, h->stream>>>(nrows, cnt, rstart);
        k_bucket<<<grid_for(L, TB), TB, 0, h->stream>>>(L, erow, ecol, rstart, cursor, tcol, tent);
        k_sort_rows<<<grid_for((int64_t)nrows * 32, TB), TB, 0, h->stream>>>(nrows, rstart, tcol, tent, scol, sent);
        k_heads<<<grid_for(nrows, TB), TB, 0, h->stream>>>(nrows, rstart, scol, head);
        h->launches += 5;
    } else {
        k_exscan<<<1, 1024, 0, h->stream>>>(nrows, cnt, rstart);
        h->launches += 1;
    }
    int Lv = 0;
    CUDA_OK(cudaMemcpyAsync(&Lv, rstart + nrows, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    k_exscan<<<1, 1024, 0, h->stream>>>(Lv, head, slotof);
    k_emit<<<grid_for(Lv > nrows ? Lv : nrows + 1, TB), TB, 0, h->stream>>>(nrows, Lv, rstart, scol, sent, esrc, head, slotof,
                                                                         out->row_ptr, out->col_idx, out->seg_ptr,
                                                                         out->seg_src);
    h->launches += 2;
    int ns = 0;
    CUDA_OK(cudaMemcpyAsync(&ns, slotof + Lv, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaGetLastError());
    out->nrows = nrows;
    out->nslots = ns;
    out->Lv = Lv;
    return 0;
}

// ---- lifecycle -------------------------------------------------------------------------
extern "C" void sqpqp_default_options(sqpqp_options* o) {
    o->rho0 = 0.1; o->sigma = 1e-6; o->alpha = 1.6;
    o->eps_abs = 1e-7; o->eps_rel = 1e-7; o->eps_inf = 1e-6;
    o->rho_eq_mult = 1e3; o->rho_min = 1e-6; o->rho_max = 1e6; o->adapt_tol = 3.0;
    o->cg_rel0 = 0.2;
    o->rb_full_mult = 1.5;
    o->polish_trigger = 5e-2; o->polish_rho = 1e4; o->polish_tol = 1e-11; o->feas_tol = 1e-9; o->dual_tol = 1e-9;
    o->max_iter = 6000; o->check_every = 25; o->ruiz_iters = 15; o->cg_max = 300; o->eig_iters = 60;
    o->polish_outer = 20; o->polish_cg_max = 3000;
    o->warm_start = 1; o->team = 0; o->threads = 0; o->smem_kb = -1; o->occupancy = 0;
    o->ipm_ic_growth = 4.0; o->ipm_ic_decay = 3.0;
    o->method = 0; o->ipm_max_iter = 200; o->fallback_max_iter = 1500; o->ipm_eps = 1e-9; o->ipm_delta0 = 1e-6; o->ipm_delta_min = 1e-8;
    o->ipm_rho0 = 1e-8; o->ipm_tau = 0.995; o->ipm_mu0 = 1.0; o->ipm_mu_min = 1e-14; o->ipm_kappa_eps = 10.0; o->ipm_refine = 0; o->verbose = 0;
}

static bool create_timing_events(sqpqp_handle h) {
    for (auto& pr : h->tev) for (auto& e : pr) if (cudaEventCreate(&e) != cudaSuccess) return false;
    for (auto& e : h->sev) if (cudaEventCreate(&e) != cudaSuccess) return false;
    h->ev0 = h->tev[0][0]; h->ev1 = h->tev[0][1];
    return true;
}
extern "C" int sqpqp_create(sqpqp_handle* out, int device) {
    if (!out) return SQPQP_E_BADARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return SQPQP_E_CUDA;
    sqpqp_handle h = new (std::nothrow) sqpqp_handle_s();
    if (!h) return SQPQP_E_NOMEM;
    h->device = device;
    memset(&h->P, 0, sizeof(Prob));
    sqpqp_default_options(&h->opts);
    DeviceGuard g(device);
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        !create_timing_events(h)) {
        delete h;
        return SQPQP_E_CUDA;
    }
    {   // side stream of the mixed-phase launch, at the highest priority: its few CTAs start as soon as any slot frees
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, hi) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) {
            delete h;
            return SQPQP_E_CUDA;
        }
    }
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
    int optin = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    h->max_dyn_smem = optin - 4096 - 1024;  // static reduction scratch + slack
    if (h->max_dyn_smem < 0) h->max_dyn_smem = 0;
    {   // shared memory per CTA: four / two CTAs per SM = (SM shared memory - n x (static scratch + 1 KB system reservation)) / n
        int per_sm = 0;
        cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
        int v = (per_sm / 4 - 4096 - 1024) & ~1023;
        h->cta4_smem = v < 16 * 1024 ? 16 * 1024 : v;
        v = (per_sm / 2 - 4096 - 1024) & ~1023;
        h->cta2_smem = v < 32 * 1024 ? 32 * 1024 : v;
    }
    cudaFuncSetAttribute(k_solve_cta<512, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_dyn_smem);
    cudaFuncSetAttribute(k_solve_cta<512, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->max_dyn_smem);
    cudaFuncSetAttribute(k_solve_cta<512, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->cta2_smem);
    cudaFuncSetAttribute(k_solve_cta<512, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->cta2_smem);
    cudaFuncSetAttribute(k_solve_cta<384, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->cta2_smem);
    cudaFuncSetAttribute(k_solve_cta<384, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->cta2_smem);
    cudaFuncSetAttribute(k_solve_cta<256, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->cta4_smem);
    cudaFuncSetAttribute(k_solve_cta<256, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->cta4_smem);
    *out = h;
    return 0;
}

static void free_problem(sqpqp_handle h) {
    cudaStreamSynchronize(h->stream);
    for (void* p : h->allocs) cudaFree(p);
    h->allocs.clear();
    memset(&h->P, 0, sizeof(Prob));
    h->setup_done = h->updated = false;
    h->d_Etrial = h->d_ftrial = nullptr;
    h->trial_valid = false;
}

extern "C" int sqpqp_destroy(sqpqp_handle h) {
    if (!h) return SQPQP_E_BADARG;
    DeviceGuard g(h->device);
    free_problem(h);
    for (auto q : h->host_owned_reg) cudaHostUnregister((void*)q);
    if (h->pin) cudaFreeHost(h->pin);
    if (h->dstage) cudaFree(h->dstage);
    for (auto& pr : h->tev) for (auto e : pr) if (e) cudaEventDestroy(e);
    for (auto e : h->sev) if (e) cudaEventDestroy(e);
    for (auto e : h->evpool) cudaEventDestroy(e);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

// Page-lock a caller-owned host buffer (cudaHostRegister) and remember its range: every later call that is handed a pointer
// inside it copies directly between that buffer and the device.  The calls stay blocking, so the buffer is free on return
// exactly as before.  A host that owns persistent arrays (the reference's sqp.dE / h_val / df / E / x / lambda ...,
// sqp.jl:16-59) registers them once after allocation.
extern "C" int sqpqp_host_register(sqpqp_handle h, void* ptr, int64_t bytes) {
    if (!h || !ptr || bytes <= 0) return SQPQP_E_BADARG;
    DeviceGuard g(h->device);
    if (is_registered(h, ptr, (size_t)bytes)) return 0;
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); e = cudaSuccess; }  // by the caller (e.g. a pinned allocator)
    else if (e != cudaSuccess) return fail_cuda(h, e, "cudaHostRegister", __LINE__);
    else h->host_owned_reg.push_back((const char*)ptr);
    h->host_pinned.emplace_back((const char*)ptr, (size_t)bytes);
    return 0;
}
extern "C" int sqpqp_host_unregister(sqpqp_handle h, void* ptr) {
    if (!h || !ptr) return SQPQP_E_BADARG;
    DeviceGuard g(h->device);
    for (size_t k = 0; k < h->host_pinned.size(); ++k)
        if (h->host_pinned[k].first == (const char*)ptr) {
            cudaStreamSynchronize(h->stream);
            for (size_t q = 0; q < h->host_owned_reg.size(); ++q)
                if (h->host_owned_reg[q] == (const char*)ptr) {
                    cudaHostUnregister(ptr);
                    h->host_owned_reg.erase(h->host_owned_reg.begin() + q);
                    break;
                }
            h->host_pinned.erase(h->host_pinned.begin() + k);
            return 0;
        }
    return fail(h, SQPQP_E_BADARG, "buffer was not registered");
}

extern "C" const char* sqpqp_last_error(sqpqp_handle h) { return h ? h->err.c_str() : "null handle"; }
extern "C" void* sqpqp_stream(sqpqp_handle h) { return h ? (void*)h->stream : nullptr; }
extern "C" int64_t sqpqp_launch_count(sqpqp_handle h) { return h ? h->launches : 0; }
// Sum of the CUDA-event durations of all solve launches of this handle so far (a step of the batched SQP may launch the QP
// phase and the restoration phase: sqpqp_last_solve_ms only sees the last one).
extern "C" double sqpqp_last_solve_ms(sqpqp_handle h);
extern "C" double sqpqp_solve_ms_total(sqpqp_handle h) {
    if (!h) return 0.0;
    sqpqp_last_solve_ms(h);
    return h->solve_ms_total;
}
extern "C" double sqpqp_last_solve_ms(sqpqp_handle h) {
    if (!h) return 0.0;
    if (h->timing_pending) {
        DeviceGuard g(h->device);
        float ms = 0.f;
        if (h->timing_is_solve) {  // every pending pair of the ring, oldest first
            while (h->tp_n > 0) {
                cudaEvent_t a = h->tev[h->tp_head][0], b = h->tev[h->tp_head][1];
                if (cudaEventSynchronize(b) == cudaSuccess && cudaEventElapsedTime(&ms, a, b) == cudaSuccess) {
                    h->last_ms = ms;
                    h->solve_ms_total += ms;
                }
                h->tp_head = (h->tp_head + 1) % kTimingRing;
                h->tp_n--;
            }
        } else if (cudaEventSynchronize(h->ev1) == cudaSuccess && cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) {
            h->last_ms = ms;
        }
        h->timing_pending = false;
    }
    return h->last_ms;
}

extern "C" const char* sqpqp_last_solve_kernel(sqpqp_handle h) { return h ? h->last_kernel.c_str() : ""; }

extern "C" int sqpqp_set_options(sqpqp_handle h, const sqpqp_options* o) {
    if (!h || !o) return SQPQP_E_BADARG;
    if (o->threads < 0 || o->threads > 512 || (o->threads % 32) != 0) return fail(h, SQPQP_E_BADARG, "threads must be a multiple of 32 in [0,512]");
    if (o->check_every < 1 || o->max_iter < 1 || !(o->rho0 > 0) || !(o->alpha > 0 && o->alpha < 2)) return fail(h, SQPQP_E_BADARG, "bad option value");
    h->opts = *o;
    return 0;
}

extern "C" int sqpqp_chol_stats(sqpqp_handle h, int64_t* nnzL, int64_t* nlev, int64_t* flops) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done) return fail(h, SQPQP_E_STATE, "setup not called");
    if (nnzL) *nnzL = h->P.has_chol ? h->chol_nnzL : 0;
    if (nlev) *nlev = h->P.has_chol ? h->chol_nlev : 0;
    if (flops) *flops = h->P.has_chol ? h->chol_flops : 0;
    return 0;
}

// Development aid (tools/gpu_prof.py): read and clear the in-kernel phase profile; all zeros unless
// the library was built with -DSQPQP_PROF.
extern "C" int sqpqp_prof_read(sqpqp_handle h, uint64_t* out32) {
    if (!h || !out32) return SQPQP_E_BADARG;
    memset(out32, 0, 32 * sizeof(uint64_t));
#ifdef SQPQP_PROF
    DeviceGuard g(h->device);
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaMemcpyFromSymbol(out32, g_prof, 32 * sizeof(uint64_t)));
    uint64_t z[32] = {0};
    CUDA_OK(cudaMemcpyToSymbol(g_prof, z, sizeof(z)));
#endif
    return 0;
}

// Development aid: copy one per-instance work array of instance b back to the host.
// kind 0: N-vector slot idx, 1: M-vector slot idx, 2: factor values, 3: solve scratch, 4: inverse diagonal, 5: wJ.
extern "C" int sqpqp_debug_read(sqpqp_handle h, int32_t kind, int32_t idx, int32_t b, double* out, int64_t count) {
    if (!h || !out) return SQPQP_E_BADARG;
    if (!h->setup_done) return fail(h, SQPQP_E_STATE, "setup not called");
    DeviceGuard g(h->device);
    Prob& P = h->P;
    if (b < 0 || b >= P.batch) return fail(h, SQPQP_E_BADARG, "instance index out of range");
    CUDA_OK(cudaStreamSynchronize(h->stream));
    const double* src = nullptr;
    size_t len = 0;
    if (kind == 0 && idx >= 0 && idx < N_COUNT) { src = P.nv[idx] + (size_t)b * P.Ne; len = P.Ne; }
    else if (kind == 1 && idx >= 0 && idx < M_COUNT) { src = P.mv[idx] + (size_t)b * P.m; len = P.m; }
    else if (kind == 2 && P.has_chol) { src = P.Lval + (size_t)b * P.chol.nnzL; len = P.chol.nnzL; }
    else if (kind == 3 && P.has_chol) { src = P.yw + (size_t)b * P.n; len = P.n; }
    else if (kind == 4 && P.has_chol) { src = P.dinv + (size_t)b * P.n; len = P.n; }
    else if (kind == 5 && P.has_chol) { src = P.wJ + (size_t)b * P.nnzJ; len = P.nnzJ; }
    else return fail(h, SQPQP_E_BADARG, "bad selector");
    if ((size_t)count < len) len = (size_t)count;
    CUDA_OK(cudaMemcpy(out, src, len * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int sqpqp_debug_set(sqpqp_handle h, int32_t what, int32_t value) {  // development knobs
    if (!h) return SQPQP_E_BADARG;
    if (what == 0 && value > 0) h->spmv_ctas_per_sm = value;
    if (what == 1) h->tail_override = value;
    if (what == 2 && (value == 0 || value == 1 || value == 2 || value == 4 || value == 8)) h->ilv_G = value;
    if (what == 3 && (value == 0 || value == 256 || value == 512 || value == 1024)) h->ilv_threads = value;
    if (what == 4 && value >= 0 && value <= 2) h->ilv_occ = value;
    if (what == 5) h->ring_enable = value != 0;
    if (what == 7) h->fuse_fwd = value != 0;
    if (what == 8) h->handoff = value;
    if (what == 9 && value >= 0 && value <= 3) h->handoff_mode = value;
    if (what == 6 && value >= 0 && value <= 2) h->ring_mode = value;
    return 0;
}

// Launch order of the batched solve: CTA slot k runs instance order[k] (a permutation of 0..batch-1; NULL = index order).
extern "C" int sqpqp_set_launch_order(sqpqp_handle h, const int32_t* order) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done) return fail(h, SQPQP_E_STATE, "setup not called");
    DeviceGuard g(h->device);
    if (!order) { h->order_user = false; return 0; }
    const size_t B = h->P.batch;
    std::vector<char> seen(B, 0);
    for (size_t k = 0; k < B; ++k) {
        if (order[k] < 0 || (size_t)order[k] >= B || seen[order[k]]) return fail(h, SQPQP_E_BADARG, "launch order is not a permutation of the batch");
        seen[order[k]] = 1;
    }
    CUDA_OK(cudaMemcpyAsync(h->d_order, order, B * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->order_user = true;
    return 0;
}

// Development aid: the interior-point loop states saved by a launch with an iteration quota (sizeof(IpmState) bytes each).
extern "C" int sqpqp_debug_read_state(sqpqp_handle h, void* out, int64_t bytes) {
    if (!h || !out) return SQPQP_E_BADARG;
    if (!h->setup_done) return fail(h, SQPQP_E_STATE, "setup not called");
    DeviceGuard g(h->device);
    const size_t need = (size_t)h->P.batch * (sizeof(IpmState) + sizeof(int));
    if ((size_t)bytes < need) return fail(h, SQPQP_E_BADARG, "state buffer too small");
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaMemcpy(out, h->P.ipm_state, (size_t)h->P.batch * sizeof(IpmState), cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy((char*)out + (size_t)h->P.batch * sizeof(IpmState), h->P.fb_flag, (size_t)h->P.batch * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int sqpqp_chol_layout(sqpqp_handle h, int64_t* tail_cols, int64_t* tree_levels) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done) return fail(h, SQPQP_E_STATE, "setup not called");
    if (tail_cols) *tail_cols = h->P.has_chol ? h->chol_tail : 0;
    if (tree_levels) *tree_levels = h->P.has_chol ? h->chol_nlev_total : 0;
    return 0;
}

extern "C" int sqpqp_num_slacks(sqpqp_handle h, int32_t* S) {
    if (!h || !S) return SQPQP_E_BADARG;
    if (!h->setup_done) return fail(h, SQPQP_E_STATE, "setup not called");
    *S = h->P.S;
    return 0;
}

// instances interleaved per CTA for a batch of this size when the caller did not choose (sqpqp_debug_set what = 2)
// Measured on BASELINE configs[4] (profiles/r02_layout_ab.md): the interleaved layout is never faster than one CTA per
// instance on the case118-shaped batch -- the G instances of a group iterate in lock step, so a group runs for the
// iteration count of its slowest member (mean 28, max 46-74 per round) -- hence auto = off.
static int ilv_auto_G(sqpqp_handle h, int batch) {
    (void)h; (void)batch;
    return 1;
}

template <int G, int NT, int MINB>
static cudaError_t launch_ilv_t(sqpqp_handle h, const DevOpts& O, int phase, const IlvDev& X, size_t dyn) {
    cudaError_t e = cudaFuncSetAttribute(k_solve_ilv<G, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
    if (e != cudaSuccess) return e;
    int grid = X.ngroups < 65535 ? X.ngroups : 65535;
    k_solve_ilv<G, NT, MINB><<<grid, NT, dyn, h->stream>>>(h->P, O, phase, X);
    return cudaGetLastError();
}
static cudaError_t launch_ilv(sqpqp_handle h, const DevOpts& O, int phase, const IlvDev& X, size_t dyn) {
    const int G = h->G, NT = h->ilv_nt, MB = h->ilv_minb;
    // the shapes kept after the A/B of profiles/r02_layout_ab.md (the layout is off by default: it lost on the headline
    // workload); other (threads, CTAs per SM) requests fall back to the nearest one
    (void)NT;
    if (G == 2) return launch_ilv_t<2, 512, 2>(h, O, phase, X, dyn);
    if (G == 4 && MB == 2) return launch_ilv_t<4, 512, 2>(h, O, phase, X, dyn);
    if (G == 4) return launch_ilv_t<4, 512, 1>(h, O, phase, X, dyn);
    if (G == 8) return launch_ilv_t<8, 512, 1>(h, O, phase, X, dyn);
    return cudaErrorInvalidConfiguration;
}

// ---- setup -------------------------------------------------------------------------------
extern "C" int sqpqp_setup_nlp(sqpqp_handle h, int32_t batch, int32_t n, int32_t m, int32_t m_lin, int64_t nnz_j,
                               const int64_t* j_row, const int64_t* j_col, int64_t nnz_h, const int64_t* h_row,
                               const int64_t* h_col, const double* x_L, const double* x_U, const double* g_L,
                               const double* g_U, int32_t bounds_per_instance) {
    if (!h) return SQPQP_E_BADARG;
    if (batch < 1 || n < 1 || m < 0 || m_lin < 0 || m_lin > m || nnz_j < 0 || nnz_h < 0) return fail(h, SQPQP_E_BADARG, "bad sizes");
    if ((nnz_j > 0 && (!j_row || !j_col)) || (nnz_h > 0 && (!h_row || !h_col)) || !x_L || !x_U || (m > 0 && (!g_L || !g_U)))
        return fail(h, SQPQP_E_BADARG, "null pointer");
    if (nnz_j + (int64_t)2 * m > std::numeric_limits<int>::max() / 4 || nnz_h > std::numeric_limits<int>::max() / 4)
        return fail(h, SQPQP_E_BADARG, "nnz too large for int32 indexing");
    for (int64_t k = 0; k < nnz_j; ++k)
        if (j_row[k] < 1 || j_row[k] > m || j_col[k] < 1 || j_col[k] > n) return fail(h, SQPQP_E_BADARG, "Jacobian COO index out of range");
    for (int64_t k = 0; k < nnz_h; ++k)
        if (h_row[k] < 1 || h_row[k] > n || h_col[k] < 1 || h_col[k] > n) return fail(h, SQPQP_E_BADARG, "Hessian COO index out of range");
    DeviceGuard g(h->device);
    free_problem(h);
    Prob& P = h->P;
    P.n = n; P.m = m; P.mlin = m_lin; P.batch = batch;
    P.has_hess = nnz_h > 0;
    h->nnzJ_coo = nnz_j;
    h->nnzH_coo = nnz_h;

    // slack columns of create_model! (subproblem_JuMP.jl:59-65): rows > m_lin get u_i, and
    // v_i when both bounds are finite; u enters with +1 except on pure <= rows (:110)
    std::vector<int> srow;
    std::vector<double> ssign;
    for (int i = m_lin; i < m; ++i) {
        bool lo = g_L[i] > -INFINITY, up = g_U[i] < INFINITY;
        srow.push_back(i);
        ssign.push_back((!lo && up) ? -1.0 : 1.0);
        if (lo && up) {
            srow.push_back(i);
            ssign.push_back(-1.0);
        }
    }
    if (bounds_per_instance)  // the slack-column structure is shared by the batch: every instance must have the finiteness
        for (int b = 1; b < batch; ++b)  // pattern of instance 0 on the nonlinear rows (a different one would silently get a wrong FR model)
            for (int i = m_lin; i < m; ++i) {
                const double l0 = g_L[i], u0 = g_U[i], lb = g_L[(size_t)b * m + i], ub = g_U[(size_t)b * m + i];
                if ((l0 > -INFINITY) != (lb > -INFINITY) || (u0 < INFINITY) != (ub < INFINITY))
                    return fail(h, SQPQP_E_BADARG, "instances of a batch must share the finite/infinite pattern of g_L, g_U on the nonlinear rows");
            }
    const int S = (int)srow.size();
    P.S = S;
    P.Ne = n + S;

    size_t stage = (size_t)(nnz_j + nnz_h) * 2 * sizeof(int64_t) + (size_t)S * 16 +
                   ((size_t)2 * n + 2 * (size_t)m) * sizeof(double) * (bounds_per_instance ? batch : 1) + 16384;
    int rc = ensure_stage(h, stage);
    if (rc) return rc;
    const int64_t* d_jr = upload(h, j_row, (size_t)nnz_j);
    const int64_t* d_jc = upload(h, j_col, (size_t)nnz_j);
    const int64_t* d_hr = upload(h, h_row, (size_t)nnz_h);
    const int64_t* d_hc = upload(h, h_col, (size_t)nnz_h);
    int* d_srow;
    double* d_ssign;
    DALLOC(d_srow, S);
    DALLOC(d_ssign, S);
    if (S) {
        CUDA_OK(cudaMemcpyAsync(d_srow, upload(h, srow.data(), (size_t)S), S * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
        CUDA_OK(cudaMemcpyAsync(d_ssign, upload(h, ssign.data(), (size_t)S), S * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
    P.slack_row = d_srow;
    P.slack_sign = d_ssign;

    // K1: three patterns
    const int LJ = (int)nnz_j + S, LH = 2 * (int)nnz_h;
    int *erow, *ecol, *esrc;
    DALLOC(erow, (LJ > LH ? LJ : LH));
    DALLOC(ecol, (LJ > LH ? LJ : LH));
    DALLOC(esrc, (LJ > LH ? LJ : LH));
    Pattern pj, pt, ph;
    const int TB = 256;
    if (LJ > 0) {
        k_entries_jac<<<grid_for(LJ, TB), TB, 0, h->stream>>>(nnz_j, d_jr, d_jc, S, d_srow, n, 0, erow, ecol, esrc);
        h->launches++;
    }
    if ((rc = build_pattern(h, LJ, erow, ecol, esrc, m, &pj))) return rc;
    if (LJ > 0) {
        k_entries_jac<<<grid_for(LJ, TB), TB, 0, h->stream>>>(nnz_j, d_jr, d_jc, S, d_srow, n, 1, erow, ecol, esrc);
        h->launches++;
    }
    if ((rc = build_pattern(h, LJ, erow, ecol, esrc, P.Ne, &pt))) return rc;
    if (LH > 0) {
        k_entries_hess<<<grid_for(nnz_h, TB), TB, 0, h->stream>>>(nnz_h, d_hr, d_hc, erow, ecol, esrc);
        h->launches++;
    }
    if ((rc = build_pattern(h, LH, erow, ecol, esrc, n, &ph))) return rc;
    int* re_n;
    DALLOC(re_n, m + 1);
    if (m > 0) {
        k_row_end_normal<<<grid_for(m, TB), TB, 0, h->stream>>>(m, pj.row_ptr, pj.col_idx, n, re_n);
        h->launches++;
    }
    P.J_rb = pj.row_ptr; P.J_re_e = pj.row_ptr + 1; P.J_re_n = re_n; P.J_col = pj.col_idx;
    P.T_rb = pt.row_ptr; P.T_col = pt.col_idx;
    P.H_rb = ph.row_ptr; P.H_col = ph.col_idx;
    P.nnzJ = pj.nslots; P.nnzT = pt.nslots; P.nnzH = ph.nslots;
    // lanes per row from the average row length of each matrix
    double nnzJn = (double)(pj.nslots - S);
    P.lgJn = lg_lanes(m ? nnzJn / m : 1.0);
    P.lgJe = lg_lanes(m ? (double)pj.nslots / m : 1.0);
    P.lgT = lg_lanes((double)pt.nslots / P.Ne);
    P.lgH = lg_lanes(n ? (double)ph.nslots / n : 1.0);

    // ---- layout of the batched interior-point path: G instances interleaved per CTA (ilv.cuh), or one CTA per instance ----
    // auto: from four instances per SM up, groups of 4 (full 32-byte sectors per gather, 2 groups resident per SM do not
    // fit its shared memory, so fewer instances than that are better served by the one-CTA-per-instance team)
    int G = h->ilv_G;
    if (G == 0) G = ilv_auto_G(h, batch);
    if (batch < 2 || m == 0) G = 1;
    h->G = G;
    h->has_ilv = h->has_ilv_fr = false;
    h->ilv_nt = 512;
    h->ilv_minb = (G == 2) ? 2 : (G == 4 && h->ilv_occ == 2 ? 2 : 1);
    // per-instance storage (workspace arrays padded to whole groups)
    const size_t B = (G > 1) ? ((size_t)batch + G - 1) / G * G : (size_t)batch;
    DALLOC(P.Jv, B * P.nnzJ); DALLOC(P.Tv, B * P.nnzT); DALLOC(P.Hv, B * P.nnzH);
    DALLOC(P.Jsv, B * P.nnzJ); DALLOC(P.Tsv, B * P.nnzT); DALLOC(P.Hsv, B * P.nnzH);
    for (int k = 0; k < N_COUNT; ++k) DALLOC(P.nv[k], B * P.Ne);
    for (int k = 0; k < M_COUNT; ++k) DALLOC(P.mv[k], B * (m > 0 ? m : 1));
    double *Jvi = nullptr, *Tvi = nullptr, *Hvi = nullptr;
    if (G > 1) { DALLOC(Jvi, B * P.nnzJ); DALLOC(Tvi, B * P.nnzT); DALLOC(Hvi, B * P.nnzH); }
    DALLOC(P.codeC, B * (m > 0 ? m : 1)); DALLOC(P.prevC, B * (m > 0 ? m : 1)); DALLOC(P.triedC, B * (m > 0 ? m : 1));
    DALLOC(P.codeB, B * P.Ne); DALLOC(P.prevB, B * P.Ne); DALLOC(P.triedB, B * P.Ne);
    DALLOC(P.rho_w, B);
    DALLOC(P.o_p, B * n); DALLOC(P.o_lam, B * (m > 0 ? m : 1)); DALLOC(P.o_mxL, B * n); DALLOC(P.o_mxU, B * n);
    DALLOC(P.o_slack, B * (S > 0 ? S : 1));
    DALLOC(P.o_info, B);
    DALLOC(P.fb_flag, B);
    DALLOC(P.ipm_state, B);
    DALLOC(h->d_order, B);
    h->order_user = false;
    DALLOC(h->d_dE, B * (size_t)nnz_j); DALLOC(h->d_hval, B * (size_t)nnz_h); DALLOC(h->d_df, B * n); DALLOC(h->d_E, B * (m > 0 ? m : 1));
    DALLOC(h->d_xk, B * n); DALLOC(h->d_delta, B); DALLOC(h->d_Eov, B * (m > 0 ? m : 1)); DALLOC(h->d_active, B);
    const size_t nb = bounds_per_instance ? B : 1;
    DALLOC(h->d_gL, nb * (m > 0 ? m : 1)); DALLOC(h->d_gU, nb * (m > 0 ? m : 1)); DALLOC(h->d_xL, nb * n); DALLOC(h->d_xU, nb * n);
    CUDA_OK(cudaMemcpyAsync(h->d_xL, upload(h, x_L, nb * n), nb * n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CUDA_OK(cudaMemcpyAsync(h->d_xU, upload(h, x_U, nb * n), nb * n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (m > 0) {
        CUDA_OK(cudaMemcpyAsync(h->d_gL, upload(h, g_L, nb * m), nb * m * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CUDA_OK(cudaMemcpyAsync(h->d_gU, upload(h, g_U, nb * m), nb * m * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
    P.gL = h->d_gL; P.gU = h->d_gU; P.xL = h->d_xL; P.xU = h->d_xU;
    P.gstride = bounds_per_instance ? m : 0;
    P.xstride = bounds_per_instance ? n : 0;
    P.df = h->d_df; P.E = h->d_E; P.Eov = nullptr; P.xk = h->d_xk; P.delta = h->d_delta; P.active = nullptr;

    // ---- symbolic Cholesky of K = P + D + J'WJ for the interior-point path (host, once) ----------
    P.has_chol = 0;
    {
        std::vector<int> hJrb(m + 1), hJre(m > 0 ? m : 1), hJc(P.nnzJ > 0 ? P.nnzJ : 1), hHrp(n + 1), hHc(P.nnzH > 0 ? P.nnzH : 1);
        CUDA_OK(cudaStreamSynchronize(h->stream));
        CUDA_OK(cudaMemcpy(hJrb.data(), pj.row_ptr, (m + 1) * sizeof(int), cudaMemcpyDeviceToHost));
        if (m) CUDA_OK(cudaMemcpy(hJre.data(), re_n, m * sizeof(int), cudaMemcpyDeviceToHost));
        if (P.nnzJ) CUDA_OK(cudaMemcpy(hJc.data(), pj.col_idx, P.nnzJ * sizeof(int), cudaMemcpyDeviceToHost));
        CUDA_OK(cudaMemcpy(hHrp.data(), ph.row_ptr, (n + 1) * sizeof(int), cudaMemcpyDeviceToHost));
        if (P.nnzH) CUDA_OK(cudaMemcpy(hHc.data(), ph.col_idx, P.nnzH * sizeof(int), cudaMemcpyDeviceToHost));
        auto up = [&](const std::vector<int>& v, const int** dst) -> int {
            int* d = nullptr;
            int rc2 = dalloc(h, &d, v.size());
            if (rc2) return rc2;
            if (!v.empty()) {
                // on the engine's own stream and waited for: the stream is non-blocking, so it does NOT order after
                // the legacy default stream a plain cudaMemcpy from pageable memory finishes its DMA on -- with the
                // 300 MB programs of a 2000-bus instance the first solve used to start on half-uploaded indices
                cudaError_t e2 = cudaMemcpyAsync(d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream);
                if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(h->stream);
                if (e2 != cudaSuccess) return fail_cuda(h, e2, "cudaMemcpy symbolic", __LINE__);
            }
            *dst = d;
            return 0;
        };
        // one large instance runs on the cooperative grid: its dense tail lives in global memory and is factorised by
        // the whole grid in panels of 32 columns (chol.cuh: dense_factor_grid)
        const bool grid_mode = batch == 1 && (size_t)P.Ne + m > 6000;
        auto upload_symbolic = [&](Symbolic& Sy, CholDev& C, int ncols, bool hasP) -> int {
            C.n = ncols; C.nnzL = Sy.nnzL; C.nlev = Sy.nlev; C.n0 = Sy.n0; C.T = Sy.T;
            C.fused_fwd = Sy.fused_fwd ? 1 : 0;
            C.Tpad = grid_mode ? ((Sy.T + GD_NB - 1) / GD_NB) * GD_NB : ((Sy.T + 3) & ~3);
            C.nphase = (int)Sy.fphase.size() / 4; C.n_aslot = (int)Sy.aslot_d.size(); C.nslotJ = (int)Sy.jrow.size();
            for (int k = 0; k < 4; ++k) Sy.fphase.push_back(0);  // the phase loop reads one entry ahead
            if (Sy.as_ab.empty()) { Sy.as_ab.push_back(0); Sy.as_ab.push_back(0); }
            if (Sy.fp_ab.empty()) { Sy.fp_ab.push_back(0); Sy.fp_ab.push_back(0); }  // gather_dot2 reads pair 0 for idle lanes
            if (Sy.ftask.empty()) Sy.ftask.assign(4, 0);
            if (Sy.aslot.empty()) { Sy.aslot.assign(4, 0); Sy.aslot_d.assign(1, -1); }
            const std::vector<int>* srcs[] = {&Sy.perm, &Sy.Lp, &Sy.Li, &Sy.Rp, &Sy.Rmid, &Sy.Rci, &Sy.lev_ptr, &Sy.fp_ab,
                                              &Sy.ftask, &Sy.fphase, &Sy.aslot, &Sy.aslot_d, &Sy.as_ab, &Sy.jrow};
            const int** dsts[] = {&C.perm, &C.Lp, &C.Li, &C.Rp, &C.Rmid, (const int**)&C.Rci, &C.lev_ptr, (const int**)&C.fp_ab,
                                  (const int**)&C.ftask, (const int**)&C.fphase, (const int**)&C.aslot,
                                  &C.aslot_d, (const int**)&C.as_ab, &C.jrow};
            for (int k = 0; k < 14; ++k) {
                int rc2 = up(*srcs[k], dsts[k]);  // cudaMalloc alignment (256 B) covers the int2 / int4 views
                if (rc2) return rc2;
            }
            // ring programs of the resident CTA team (chol.cuh): chunk images for RING_S stages of kRingStageBytes; strips of
            // up to 512 slots (one per thread of the resident launch), at most 4 pairs per lane
            C.ring_ok = 0; C.rprog = nullptr; C.ring_nL = 0; C.ring_stage_words = 0;
            if (!grid_mode && G == 1 && h->ring_enable) {
                RingProg R;
                build_ring_program(Sy, hasP, 512, 4, kRingStageBytes, RING_S, R);
                if (R.ok) {
                    int rc2 = up(R.words, &C.rprog);
                    if (rc2) return rc2;
                    C.ring_ok = 1; C.ring_nL = R.nL; C.ring_stage_words = (R.stage_words + 3) & ~3;
                    for (int sg = 0; sg < 3; ++sg) {
                        C.rseg_n[sg] = R.seg_count[sg];
                        for (int q = 0; q < RING_S; ++q) {
                            const bool on = q < R.seg_count[sg];
                            C.rseg_off[sg][q] = on ? R.chunk_off[R.seg_first[sg] + q] : -1;
                            C.rseg_bytes[sg][q] = on ? 4 * R.chunk_len[R.seg_first[sg] + q] : 0;
                        }
                    }
                }
            }
            return 0;
        };
        // Dense tail of the factor (chol.cuh): as many top levels of the elimination tree as fit the
        // shared memory one CTA gets when four CTAs share an SM, next to the inverse diagonal and the
        // solve scratch when those fit as well.  A single large instance runs on the cooperative grid
        // (no CTA-local shared memory across the team) and keeps the plain level-scheduled code.
        // shared-memory budget of one interleaved CTA: the opt-in maximum (one CTA per SM) or half an SM, minus its
        // static reduction scratch
        const size_t ilv_static = (size_t)2 * ILV_KMAX * (h->ilv_nt / 32) * G * sizeof(double);
        int per_sm_smem = 0, optin_smem = 0;
        cudaDeviceGetAttribute(&per_sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, h->device);
        cudaDeviceGetAttribute(&optin_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
        const size_t ilv_budget = (h->ilv_minb >= 2 ? (size_t)per_sm_smem / 2 - 1024 : (size_t)optin_smem) - ilv_static - 256;
        auto ilv_yw_resident = [&](int ncols) -> bool {  // solve scratch next to a tail of at least 48 columns?
            return (size_t)ncols * G * sizeof(double) + (size_t)52 * 53 / 2 * G * sizeof(double) <= ilv_budget;
        };
        auto tail_cap = [&](int ncols) -> int {
            // grid team: the chain at the top of the tree (~740 columns on the 2000-bus network) becomes one dense block.
            // Measured on that network with the forward sweep fused into the factor phases (ms per interior-point iteration by
            // cap): 384: 5.12, 512: 2.97, 640: 1.81, 768: 1.51, 896: 1.56, 1024: 1.69 -- 765 columns = 2.3 MB packed (L2-resident)
            if (grid_mode) return h->tail_override >= 0 ? h->tail_override : 768;
            if (G > 1) {
                size_t cap = ilv_budget / sizeof(double) / G;  // doubles per instance
                if (ilv_yw_resident(ncols)) cap -= ncols;
                int lim = h->tail_override >= 0 ? h->tail_override : 96;
                int t = 0;
                while (t < lim && (size_t)(t + 4) * (t + 5) / 2 <= cap) ++t;
                return t;
            }
            if (h->tail_override >= 0) return h->tail_override;  // sqpqp_debug_set(h, 1, columns): tuning runs only
            size_t cap = (size_t)h->cta2_smem / sizeof(double);
            size_t vec = 2 * (size_t)((ncols + 1) & ~1);
            if (vec <= cap / 2) cap -= vec;
            int t = 0;
            // 96 columns: beyond that the dense factorisation of the (only ~40 % full) tail costs more than the
            // sparse levels it replaces (measured on the case118-shaped batch: T = 48 / 92 / 128 -> 79 / 61 / 67 ms)
            while (t < 96 && (size_t)(t + 4) * (t + 5) / 2 + (size_t)((t + 2) & ~1) <= cap) ++t;
            return t;
        };
        // slot lists for 32 / G lanes per task + shared-memory placement of the interleaved path
        auto build_ilv = [&](const Symbolic& Sy, const CholDev& base, int ncols, IlvDev* X, size_t* dyn) -> int {
            int lgG = 0;
            while ((1 << lgG) < G) ++lgG;
            SlotProg sp;
            build_slot_programs(Sy, 5 - lgG, h->ilv_nt / G, sp);
            X->C = base;
            X->C.fused_fwd = 0;  // its own (unfused) slot lists
            X->C.nphase = (int)sp.fphase.size() / 4;
            X->C.n_aslot = (int)sp.aslot_d.size();
            for (int k = 0; k < 4; ++k) sp.fphase.push_back(0);
            if (sp.ftask.empty()) sp.ftask.assign(4, 0);
            if (sp.aslot.empty()) { sp.aslot.assign(4, 0); sp.aslot_d.assign(1, -1); }
            int rc3;
            if ((rc3 = up(sp.ftask, (const int**)&X->C.ftask)) || (rc3 = up(sp.fphase, (const int**)&X->C.fphase)) ||
                (rc3 = up(sp.aslot, (const int**)&X->C.aslot)) || (rc3 = up(sp.aslot_d, &X->C.aslot_d)))
                return rc3;
            X->Jvi = Jvi; X->Tvi = Tvi; X->Hvi = Hvi;
            X->ngroups = (int)(B / G);
            const size_t Tp = ((size_t)Sy.T + 3) & ~(size_t)3;
            size_t off = 0;
            X->off_D = -1; X->off_yw = -1; X->off_dinv = -1;
            if (Sy.T > 0) { X->off_D = 0; off = Tp * (Tp + 1) / 2 * G; }
            if ((off + (size_t)ncols * G) * sizeof(double) <= ilv_budget) { X->off_yw = (int)off; off += (size_t)ncols * G; }
            *dyn = off * sizeof(double);
            if (*dyn > ilv_budget) return fail(h, SQPQP_E_STATE, "dense tail of the factor does not fit the shared memory of the interleaved launch");
            return 0;
        };
        CUDA_OK(cudaStreamSynchronize(h->stream));
        {   // CSR-stream SpMV plans (spmv.cuh): row blocks of <= SPMV_CHUNK value slots, shared by the batch
            std::vector<int> hTrp(n + 1), blk;
            CUDA_OK(cudaMemcpy(hTrp.data(), pt.row_ptr, (n + 1) * sizeof(int), cudaMemcpyDeviceToHost));
            const int* d = nullptr;
            int rc2;
            spmv_blocks(m, hJrb.data(), hJre.data(), blk);
            if ((rc2 = up(blk, &d))) return rc2;
            h->planJ = SpmvPlan{(int)blk.size() - 1, d, pj.row_ptr, re_n, pj.col_idx, n, m, P.nnzJ};
            spmv_blocks(n, hTrp.data(), hTrp.data() + 1, blk);
            if ((rc2 = up(blk, &d))) return rc2;
            h->planT = SpmvPlan{(int)blk.size() - 1, d, pt.row_ptr, pt.row_ptr + 1, pt.col_idx, m, n, P.nnzT};
            spmv_blocks(n, hHrp.data(), hHrp.data() + 1, blk);
            if ((rc2 = up(blk, &d))) return rc2;
            h->planH = SpmvPlan{(int)blk.size() - 1, d, ph.row_ptr, ph.row_ptr + 1, ph.col_idx, n, n, P.nnzH};
        }
        DALLOC(P.wJ, B * (size_t)(P.nnzJ > 0 ? P.nnzJ : 1));
        Symbolic Sy = symbolic_analyze(n, m, hJrb.data(), hJre.data(), hJc.data(), hHrp.data(), hHc.data(), 512, tail_cap(n), h->fuse_fwd != 0);
        if (Sy.ok && (int64_t)Sy.fp_ab.size() < ((int64_t)1 << 29) && Sy.nnzL < (1 << 26)) {
            int rc2 = upload_symbolic(Sy, P.chol, n, true);
            if (rc2) return rc2;
            DALLOC(P.Lval, B * (size_t)Sy.nnzL);
            DALLOC(P.yw, B * (size_t)n);
            DALLOC(P.dinv, B * (size_t)n);
            if (grid_mode && Sy.T > 0) DALLOC(P.Dtail, (size_t)P.chol.Tpad * (P.chol.Tpad + 1) / 2);
            P.has_chol = 1;
            h->chol_nnzL = Sy.nnzL; h->chol_nlev = Sy.nlev; h->chol_flops = Sy.flops; h->chol_tail = Sy.T;
            h->chol_nlev_total = Sy.nlev_total;
            if (G > 1) {
                rc2 = build_ilv(Sy, P.chol, n, &h->ilv, &h->ilv_dyn);
                if (rc2) return rc2;
                h->has_ilv = true;
            }
        }
        // feasibility-restoration LP: columns [J | S], no quadratic term
        P.has_chol_fr = 0;
        if (S > 0 && m > 0) {
            Symbolic Sf = symbolic_analyze(P.Ne, m, hJrb.data(), hJrb.data() + 1, hJc.data(), nullptr, nullptr, 512, tail_cap(P.Ne), h->fuse_fwd != 0);
            if (Sf.ok && (int64_t)Sf.fp_ab.size() < ((int64_t)1 << 29) && Sf.nnzL < (1 << 26)) {
                int rc2 = upload_symbolic(Sf, P.chol_fr, P.Ne, false);
                if (rc2) return rc2;
                DALLOC(P.Lval_fr, B * (size_t)Sf.nnzL);
                DALLOC(P.yw_fr, B * (size_t)P.Ne);
                DALLOC(P.dinv_fr, B * (size_t)P.Ne);
                if (grid_mode && Sf.T > 0) DALLOC(P.Dtail_fr, (size_t)P.chol_fr.Tpad * (P.chol_fr.Tpad + 1) / 2);
                P.has_chol_fr = 1;
                if (G > 1) {
                    rc2 = build_ilv(Sf, P.chol_fr, P.Ne, &h->ilv_fr, &h->ilv_fr_dyn);
                    if (rc2) return rc2;
                    h->has_ilv_fr = true;
                }
            }
        }
    }

    // cooperative-grid scratch
    int bps = 0;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_solve_grid, 256, 0));
    if (bps < 1) bps = 1;
    if (bps > 2) bps = 2;
    h->coop_blocks = bps * h->num_sms;
    P.gred_stride = h->coop_blocks;
    DALLOC(P.gred, (size_t)2 * SQPQP_MAX_RED * P.gred_stride);

    // scatter jobs
    h->jobJ = ScatterJob{P.nnzJ, (int)nnz_j, pj.seg_ptr, pj.seg_src, h->d_dE, d_ssign, P.Jv, Jvi, G};
    h->jobT = ScatterJob{P.nnzT, (int)nnz_j, pt.seg_ptr, pt.seg_src, h->d_dE, d_ssign, P.Tv, Tvi, G};
    h->jobH = ScatterJob{P.nnzH, (int)nnz_h, ph.seg_ptr, ph.seg_src, h->d_hval, nullptr, P.Hv, Hvi, G};
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaGetLastError());
    h->setup_done = true;
    h->generic = false;
    return 0;
}

// ---- update ------------------------------------------------------------------------------
static int launch_scatter(sqpqp_handle h, const double* dE, const double* hval) {
    ScatterJob a = h->jobJ, b = h->jobT, c = h->jobH;
    a.vals = dE; b.vals = dE; c.vals = hval;
    int64_t work = (int64_t)(a.nslots > b.nslots ? a.nslots : b.nslots) * h->P.batch;
    k_scatter<<<grid_for(work, 256), 256, 0, h->stream>>>(a, b, c, h->P.batch);
    h->launches++;
    return 0;
}

extern "C" int sqpqp_update_nlp(sqpqp_handle h, const double* dE, const double* h_val, const double* df, const double* E) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done) return fail(h, SQPQP_E_STATE, "setup not called");
    Prob& P = h->P;
    if ((h->nnzJ_coo && !dE) || !df || (P.m && !E) || (h->nnzH_coo && !h_val)) return fail(h, SQPQP_E_BADARG, "null pointer");
    DeviceGuard g(h->device);
    const size_t B = P.batch;
    size_t bytes = B * ((size_t)h->nnzJ_coo + h->nnzH_coo + P.n + P.m) * sizeof(double) + 8192;
    int rc = ensure_stage(h, bytes);
    if (rc) return rc;
    // pinned staging -> async H2D straight into the owned device arrays
    struct Cp { double* dst; const double* src; size_t cnt; } cps[4] = {
        {h->d_dE, dE, B * (size_t)h->nnzJ_coo}, {h->d_hval, h_val, B * (size_t)h->nnzH_coo}, {h->d_df, df, B * P.n}, {h->d_E, E, B * P.m}};
    for (auto& c : cps) {
        if (!c.cnt) continue;
        size_t off = h->stage_off;
        h->stage_off = align256(off + c.cnt * sizeof(double));
        CUDA_OK(stage_h2d(h, c.dst, off, c.src, c.cnt * sizeof(double)));
    }
    P.df = h->d_df; P.E = h->d_E;
    launch_scatter(h, h->d_dE, h->d_hval);
    // the pinned buffer is reused by the next call: wait for the copies (not for the scatter's consumers)
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaGetLastError());
    h->updated = true;
    return 0;
}

extern "C" int sqpqp_update_nlp_device(sqpqp_handle h, const double* dE, const double* h_val, const double* df, const double* E) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done) return fail(h, SQPQP_E_STATE, "setup not called");
    Prob& P = h->P;
    if ((h->nnzJ_coo && !dE) || !df || (P.m && !E) || (h->nnzH_coo && !h_val)) return fail(h, SQPQP_E_BADARG, "null pointer");
    DeviceGuard g(h->device);
    P.df = df; P.E = E;
    launch_scatter(h, dE, h_val);
    CUDA_OK(cudaGetLastError());
    h->updated = true;
    return 0;
}

// ---- device-side ACOPF evaluator (acopf.cuh) ---------------------------------------------------
extern "C" int sqpqp_acopf_setup(sqpqp_handle h, int32_t nb, int32_t ng, int32_t nl, int32_t ref_bus, const int32_t* f_bus,
                                 const int32_t* t_bus, const int32_t* gen_bus, const double* oa, const double* oc, const double* os,
                                 const double* cost2, const double* cost1, const double* cost0, const double* gs, const double* bs,
                                 int32_t nbal, const int32_t* bal_ptr, const int32_t* bal_col, const int32_t* bal_kind,
                                 const double* bal_const, int32_t nsh, const int32_t* sh_bus) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done) return fail(h, SQPQP_E_STATE, "setup not called");
    Prob& P = h->P;
    if (nb < 1 || ng < 0 || nl < 0 || ref_bus < 0 || ref_bus >= nb) return fail(h, SQPQP_E_BADARG, "bad network size");
    // the evaluator writes the COO values in the order fixed by the layout documented in acopf.cuh: check the sizes
    if (P.n != 2 * nb + 2 * ng + 4 * nl || P.m != 1 + 2 * nb + 8 * nl || h->nnzJ_coo != (int64_t)8 * nl + 1 + nbal + 20 * nl ||
        h->nnzH_coo != (int64_t)ng + 4 * nl + 2 * nsh + 36 * nl)
        return fail(h, SQPQP_E_BADARG, "network does not match the NLP given to sqpqp_setup_nlp (n, m, COO lengths)");
    DeviceGuard g(h->device);
    AcopfDev& A = h->acopf;
    auto upi = [&](const int32_t* src, size_t cnt, const int** dst) -> int {
        int* d = nullptr;
        int rc = dalloc(h, &d, cnt);
        if (rc) return rc;
        if (cnt) CUDA_OK(cudaMemcpyAsync(d, src, cnt * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CUDA_OK(cudaStreamSynchronize(h->stream));
        *dst = d;
        return 0;
    };
    auto upd = [&](const double* src, size_t cnt, const double** dst) -> int {
        double* d = nullptr;
        int rc = dalloc(h, &d, cnt);
        if (rc) return rc;
        if (cnt) CUDA_OK(cudaMemcpyAsync(d, src, cnt * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        CUDA_OK(cudaStreamSynchronize(h->stream));
        *dst = d;
        return 0;
    };
    int rc;
    if ((rc = upi(f_bus, nl, &A.f_bus)) || (rc = upi(t_bus, nl, &A.t_bus)) || (rc = upi(gen_bus, ng, &A.gen_bus)) ||
        (rc = upd(oa, (size_t)4 * nl, &A.oa)) || (rc = upd(oc, (size_t)4 * nl, &A.oc)) || (rc = upd(os, (size_t)4 * nl, &A.os)) ||
        (rc = upd(cost2, ng, &A.cost2)) || (rc = upd(cost1, ng, &A.cost1)) || (rc = upd(cost0, ng, &A.cost0)) ||
        (rc = upd(gs, nb, &A.gs)) || (rc = upd(bs, nb, &A.bs)) || (rc = upi(bal_ptr, (size_t)2 * nb + 1, &A.bal_ptr)) ||
        (rc = upi(bal_col, nbal, &A.bal_col)) || (rc = upi(bal_kind, nbal, &A.bal_kind)) || (rc = upd(bal_const, nbal, &A.bal_const)) ||
        (rc = upi(sh_bus, nsh, &A.sh_bus)))
        return rc;
    DALLOC(h->d_f, P.batch);
    A.nb = nb; A.ng = ng; A.nl = nl; A.ref_bus = ref_bus; A.nbal = nbal; A.nsh = nsh;
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

// eval_functions! (sqp.jl:86-104) on the device for the masked instances + the value scatter; f, E and grad f come back
extern "C" int sqpqp_acopf_eval_update(sqpqp_handle h, const double* x, const double* lambda, const int32_t* mask, double* f,
                                       double* E, double* df) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done || h->acopf.nb == 0) return fail(h, SQPQP_E_STATE, "setup / acopf_setup not called");
    if (!x || !lambda) return fail(h, SQPQP_E_BADARG, "null pointer");
    DeviceGuard g(h->device);
    Prob& P = h->P;
    const size_t B = P.batch;
    int rc = ensure_stage(h, B * ((size_t)3 * P.n + 3 * (size_t)P.m + 32) * sizeof(double) + B * 16 + 16384);
    if (rc) return rc;
    AcopfArgs G;
    G.x = upload(h, x, B * P.n);
    G.lam = upload(h, lambda, B * P.m);
    G.mask = mask ? upload(h, mask, B) : nullptr;
    G.dE = h->d_dE; G.hval = h->d_hval; G.df = h->d_df; G.E = h->d_E; G.f = h->d_f;
    G.nnzJ = (int)h->nnzJ_coo; G.nnzH = (int)h->nnzH_coo;
    G.fonly = 0;
    k_acopf_eval<<<(int)(B < 65535 ? B : 65535), 256, 0, h->stream>>>(h->acopf, G, P.n, P.m, (int)B);
    h->launches++;
    P.df = h->d_df; P.E = h->d_E;
    launch_scatter(h, h->d_dE, h->d_hval);
    download(h, h->d_f, f, B);
    download(h, h->d_E, E, B * P.m);
    download(h, h->d_df, df, B * P.n);
    h->updated = true;
    return finish(h);
}

// f and g at a TRIAL point (compute_phi with alpha > 0, sqp.jl:170-183: x + p of do_step!, sqp_trust_region.jl:515-530) on the
// device: function values only, into trial buffers that a following sqpqp_merit(..., E_trial = NULL, f_trial = NULL) reads in
// place.  Instances with mask == 0 get a copy of the current E and f (their trial values are not used).  f / E may be NULL
// (nothing comes back to the host).
extern "C" int sqpqp_acopf_eval_trial(sqpqp_handle h, const double* x_trial, const int32_t* mask, double* f, double* E) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done || h->acopf.nb == 0) return fail(h, SQPQP_E_STATE, "setup / acopf_setup not called");
    if (!x_trial) return fail(h, SQPQP_E_BADARG, "null pointer");
    DeviceGuard g(h->device);
    Prob& P = h->P;
    const size_t B = P.batch;
    if (!h->d_Etrial) {
        DALLOC(h->d_Etrial, B * (size_t)(P.m > 0 ? P.m : 1));
        DALLOC(h->d_ftrial, B);
    }
    int rc = ensure_stage(h, B * ((size_t)P.n + P.m + 32) * sizeof(double) + B * 16 + 16384);
    if (rc) return rc;
    CUDA_OK(cudaMemcpyAsync(h->d_Etrial, h->d_E, B * P.m * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CUDA_OK(cudaMemcpyAsync(h->d_ftrial, h->d_f, B * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    AcopfArgs G;
    G.x = upload(h, x_trial, B * P.n);
    G.lam = nullptr;
    G.mask = mask ? upload(h, mask, B) : nullptr;
    G.dE = G.hval = G.df = nullptr; G.E = h->d_Etrial; G.f = h->d_ftrial;
    G.nnzJ = (int)h->nnzJ_coo; G.nnzH = (int)h->nnzH_coo;
    G.fonly = 1;
    k_acopf_eval<<<(int)(B < 65535 ? B : 65535), 256, 0, h->stream>>>(h->acopf, G, P.n, P.m, (int)B);
    h->launches++;
    h->trial_valid = true;
    if (f) download(h, h->d_ftrial, f, B);
    if (E) download(h, h->d_Etrial, E, B * P.m);
    return finish(h);
}

// ---- solve -------------------------------------------------------------------------------
// Greedy shared-memory placement, hottest arrays first: the PCG working set (6 N-vectors +
// 2 M-vectors), then the scaled matrix values (read 3x per PCG iteration), then the ADMM
// iterates, then the polish scratch.  Warm-start vectors must survive the kernel and stay global.
static void place_arrays(const Prob& P, int phase, size_t budget_bytes, bool ipm, bool vectors, Placement* pl, bool ring = false) {
    const int N = (phase == SQPQP_PHASE_FR) ? P.Ne : P.n, M = P.m > 0 ? P.m : 1;
    for (int k = 0; k < N_COUNT; ++k) pl->n_off[k] = -1;
    for (int k = 0; k < M_COUNT; ++k) pl->m_off[k] = -1;
    pl->jsv = pl->tsv = pl->hsv = pl->lval = pl->yw = pl->dtail = pl->dcol = pl->dinv = pl->ring = -1;
    pl->vec_resident = 0;
    size_t cap = budget_bytes / sizeof(double), off = 0;
    auto take = [&](int* slot, size_t len) {
        len = (len + 1) & ~(size_t)1;  // keep 16-byte alignment
        if (off + len <= cap) { *slot = (int)off; off += len; }
    };
    if (ipm) {  // interior point: the dense tail of the factor (lives only here), then the solve scratch
        const CholDev& C = (phase == SQPQP_PHASE_FR) ? P.chol_fr : P.chol;
        if (C.T > 0) {
            const size_t Tp = ((size_t)C.T + 3) & ~(size_t)3;  // dense_factor works in panels of 4 columns
            take(&pl->dtail, Tp * (Tp + 1) / 2);
            take(&pl->dcol, C.T);
        }
        take(&pl->dinv, C.n);
        take(&pl->yw, C.n);
        if (ring && C.ring_ok) {
            // the ring stages, then only the part of L the ring programs touch (the tail block lives in D alone), then the
            // vectors the sparse products gather from: with the indices streamed, those are the L2 round trips left
            take(&pl->ring, ((size_t)RING_S * C.ring_stage_words * sizeof(int) + 7) / 8);
            if (pl->ring >= 0) {
                take(&pl->lval, (size_t)C.ring_nL + 1);  // + the zero entry
                take(&pl->n_off[N_X], N);
                take(&pl->m_off[M_T], M);
                take(&pl->n_off[N_XT], N);
            }
        }
        if (vectors && pl->lval < 0) take(&pl->lval, C.nnzL);
    }
    if (vectors) {
        size_t before = (pl->ring >= 0) ? 0 : off;
        const int hotN[] = {N_P, N_KP, N_R, N_XT, N_MINV, N_DSH};
        for (int k : hotN) if (pl->n_off[k] < 0) take(&pl->n_off[k], N);
        if (pl->m_off[M_T] < 0) take(&pl->m_off[M_T], M);
        take(&pl->m_off[M_RC], M);
        take(&pl->tsv, P.nnzT);
        take(&pl->jsv, P.nnzJ);
        if (phase == SQPQP_PHASE_QP || phase == SQPQP_PHASE_SOC) take(&pl->hsv, P.nnzH);
        const int warmN[] = {N_X, N_ZB, N_YB, N_RB, N_Q, N_XL, N_XU, N_HD, N_D};
        const int warmM[] = {M_ZC, M_YC, M_RL, M_RU, M_ES, M_AX};
        for (int k : warmN) if (pl->n_off[k] < 0) take(&pl->n_off[k], N);
        for (int k : warmM) take(&pl->m_off[k], M);
        const int coldN[] = {N_MASK, N_XFIX, N_TMP, N_TMP2};
        const int coldM[] = {M_RW, M_BC, M_YP, M_TMP};
        for (int k : coldN) take(&pl->n_off[k], N);
        for (int k : coldM) take(&pl->m_off[k], M);
        pl->vec_resident = off > before;
    }
    pl->total = (int)off;
}

static int pick_threads(sqpqp_handle h, int phase) {
    if (h->opts.threads) return h->opts.threads;
    int len = h->P.m > h->P.n ? h->P.m : h->P.n;
    if (phase == SQPQP_PHASE_FR) len = h->P.m > h->P.Ne ? h->P.m : h->P.Ne;
    int t = 64;
    while (t < 512 && t * 2 <= len) t *= 2;
    return t;
}

// Launch order of the second stage of a two-stage batched solve: the instances the iteration quota stopped (flag 3), by
// descending predicted remaining work; finished instances (their CTAs return at once) behind them.  Rank sort: B is a few
// thousand at most, every thread counts the keys ahead of its own through shared-memory tiles.
__device__ __forceinline__ double unfinished_key(const IpmState& s, int flag, double rho0) {
    if (flag != 3) return -1.0;
    // failed factorisations so far (inertia corrections: the subproblem is indefinite where the iterates are), a shift still
    // above its floor, and how far the barrier parameter still has to fall
    const double fails = (double)(s.nfact - s.it);
    const double mu = log10(fmax(s.mu_t, 1e-12)) + 12.0;
    return 1.0 + 100.0 * fails + (s.rho_p > rho0 ? 50.0 : 0.0) + 2.0 * mu;
}
__global__ void __launch_bounds__(256) k_rank_unfinished(const IpmState* __restrict__ st, const int* __restrict__ flag, int* __restrict__ order,
                                                         int B, double rho0) {
    __shared__ double tile[256];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const double ki = i < B ? unfinished_key(st[i], flag[i], rho0) : 0.0;
    int rank = 0;
    for (int j0 = 0; j0 < B; j0 += 256) {
        const int j = j0 + threadIdx.x;
        tile[threadIdx.x] = j < B ? unfinished_key(st[j], flag[j], rho0) : -2.0;
        __syncthreads();
        const int lim = min(256, B - j0);
        for (int t = 0; t < lim; ++t) {
            const double kj = tile[t];
            rank += (kj > ki || (kj == ki && j0 + t < i)) ? 1 : 0;
        }
        __syncthreads();
    }
    if (i < B) order[rank] = i;
}

static int solve_team(sqpqp_handle h, int phase) {
    const Prob& P = h->P;
    int team = h->opts.team;
    if (team == 0) team = (P.batch == 1 && (size_t)P.Ne + P.m > 6000) ? 2 : 1;
    // a dense tail laid out for the CTA team (panels of 4, shared memory) cannot be run by the grid team and vice versa
    if (team == 2 && ((phase == SQPQP_PHASE_FR ? P.chol_fr.T : P.chol.T) > 0) && !(phase == SQPQP_PHASE_FR ? P.Dtail_fr : P.Dtail)) team = 1;
    if (team == 1 && ((phase == SQPQP_PHASE_FR ? P.chol_fr.T : P.chol.T) > 0) && (phase == SQPQP_PHASE_FR ? P.Dtail_fr : P.Dtail)) team = 2;
    return team;
}

// event pair of a solve launch: a small ring, so that back-to-back launches need no host synchronisation between them; a slot
// is only waited for when the ring wraps around to it
static int begin_solve_timing(sqpqp_handle h) {
    if (h->timing_pending && h->timing_is_solve && h->tp_n == kTimingRing) sqpqp_last_solve_ms(h);
    if (!h->timing_is_solve && h->timing_pending) sqpqp_last_solve_ms(h);
    {
        const int slot = (h->tp_head + h->tp_n) % kTimingRing;
        h->ev0 = h->tev[slot][0]; h->ev1 = h->tev[slot][1];
    }
    CUDA_OK(cudaEventRecord(h->ev0, h->stream));
    return 0;
}
static int end_solve_timing(sqpqp_handle h) {
    CUDA_OK(cudaEventRecord(h->ev1, h->stream));
    h->timing_pending = true;
    h->timing_is_solve = true;
    h->tp_n++;
    return 0;
}

// the kernels of ONE phase (interior-point launch, hand-off launch, masked ADMM launch) on stream st
static int launch_solve_kernels(sqpqp_handle h, int phase, cudaStream_t st) {
    Prob& P = h->P;
    const size_t B = P.batch;
    DevOpts O{h->opts, 0, 0};
    const int team = solve_team(h, phase);
    if (team == 2) {
        void* args[] = {(void*)&P, (void*)&O, (void*)&phase};
        CUDA_OK(cudaLaunchCooperativeKernel((void*)k_solve_grid, dim3(h->coop_blocks), dim3(256), args, 0, st));
        h->last_kernel = "k_solve_grid";
    } else {
        int threads = pick_threads(h, phase);
        int grid = (int)(B < 65535 ? B : 65535);
        // Shared memory: the factorisation's dense tail, inverse diagonal and solve scratch always; the
        // work vectors and matrix values only when one CTA owns the SM (small batches).  Large batches:
        // several CTAs share an SM and hide each other's latency, and the shared index programs are
        // re-read through L1, so L1 capacity beats vector residency (measured, profiles/r01_tuning.md).
        // more instances than SMs: two CTAs per SM, so that up to 2 x num_sms instances are co-resident in ONE wave (with
        // one CTA per SM a shard of 149..295 instances ran as one full wave plus a partial one: the N = 4 cliff of round 1)
        const bool many = B > (size_t)h->num_sms;
        int occ = h->opts.occupancy;  // 0 auto
        // measured (profiles/r01_tuning.md): two CTAs per SM beat four 256-thread ones (shorter per-instance latency for
        // the stragglers of a batch) and one 1024-thread CTA (too little work per phase); 384 threads (80 registers,
        // a third of the spills) beat 512 (64 registers) by 3 %
        if (occ == 0) occ = many ? 2 : 1;
        if (!h->opts.threads && many && threads > 256 && occ != 2) threads = 256;
        if (!h->opts.threads && occ == 2) threads = 384;
        bool ipm = (phase == SQPQP_PHASE_FR ? P.has_chol_fr : P.has_chol) && h->opts.method != 1;
        const CholDev& CD = (phase == SQPQP_PHASE_FR) ? P.chol_fr : P.chol;
        size_t budget = occ >= 8 ? 24 * 1024 : (occ >= 3 ? (size_t)h->cta4_smem : (occ == 2 ? (size_t)h->cta2_smem : (size_t)h->max_dyn_smem));
        bool vectors = occ < 2;
        if (h->opts.smem_kb >= 0) {  // explicit budget for the vectors (0 = none)
            vectors = h->opts.smem_kb > 0;
            if ((size_t)h->opts.smem_kb * 1024 < budget && vectors) budget = (size_t)h->opts.smem_kb * 1024;
        }
        Placement pl;
        // the resident launch (one CTA per SM) streams its index programs through the shared-memory ring when they were built
        const bool want_ring = ipm && occ == 1 && vectors && CD.ring_ok && h->ring_mode != 1 && !(h->G > 1);
        place_arrays(P, phase, budget, ipm, vectors, &pl, want_ring);
        if (want_ring && (pl.ring < 0 || pl.lval < 0 || pl.dinv < 0 || pl.yw < 0 || (CD.T > 0 && pl.dtail < 0)))
            place_arrays(P, phase, budget, ipm, vectors, &pl, false);  // does not fit next to the factor: slot lists from L2
        if (ipm && CD.T > 0 && pl.dtail < 0 && occ >= 3) {  // the tail was sized for two CTAs per SM
            occ = 2;
            if (!h->opts.threads) threads = 384;
            budget = (size_t)h->cta2_smem;
            place_arrays(P, phase, budget, ipm, vectors, &pl);
        }
        if (ipm && CD.T > 0 && pl.dtail < 0) return fail(h, SQPQP_E_STATE, "dense tail of the factor does not fit the shared-memory budget of this launch configuration");
        size_t dyn = (size_t)pl.total * sizeof(double);
        // interior-point launch, then the ADMM launch for the instances it flagged (a no-op for the others); with
        // options.method == 1 only the ADMM launch (all instances), with method == 2 only the interior-point one
        const int cfg = (occ >= 3 && threads <= 256) ? 4 : (occ >= 2 ? 2 : 1);
        auto launch = [&](int mode, const DevOpts& OO) {
            if (cfg == 4) {
                if (mode == 1) k_solve_cta<256, 4, 1><<<grid, threads, dyn, st>>>(P, OO, phase, pl);
                else k_solve_cta<256, 4, 2><<<grid, threads, dyn, st>>>(P, OO, phase, pl);
            } else if (cfg == 2 && threads <= 384) {  // 85 registers per thread instead of 64
                if (mode == 1) k_solve_cta<384, 2, 1><<<grid, threads, dyn, st>>>(P, OO, phase, pl);
                else k_solve_cta<384, 2, 2><<<grid, threads, dyn, st>>>(P, OO, phase, pl);
            } else if (cfg == 2) {
                if (mode == 1) k_solve_cta<512, 2, 1><<<grid, threads, dyn, st>>>(P, OO, phase, pl);
                else k_solve_cta<512, 2, 2><<<grid, threads, dyn, st>>>(P, OO, phase, pl);
            } else {
                if (mode == 1) k_solve_cta<512, 1, 1><<<grid, threads, dyn, st>>>(P, OO, phase, pl);
                else k_solve_cta<512, 1, 2><<<grid, threads, dyn, st>>>(P, OO, phase, pl);
            }
            h->launches++;
        };
        if (threads > 512) threads = 512;
        if (cfg == 4 && threads > 256) threads = 256;
        h->last_ring = pl.ring >= 0;
        // Hand-off: the throughput launch (several CTAs per SM, values in L2, ~0.49 ms per interior-point iteration) stops an
        // instance after `quota` iterations; the stragglers -- the few instances of a batch that need 2-4 x the mean -- are
        // continued by a second, resident launch (one CTA per SM, ring, ~0.34 ms per iteration) that every other CTA leaves at once.
        Placement pl2;
        size_t dyn2 = 0;
        int quota = 0;
        const int hmode = h->handoff_mode;
        if (ipm && hmode != 0 && !(h->G > 1) && h->opts.method != 1 && h->handoff > 0 && h->handoff < h->opts.ipm_max_iter) quota = h->handoff;
        if (hmode == 0 && ipm && cfg >= 2 && !(h->G > 1) && CD.ring_ok && h->ring_mode != 1 && h->opts.method != 1 && h->handoff != 0) {
            place_arrays(P, phase, (size_t)h->max_dyn_smem, true, true, &pl2, true);
            if (pl2.ring >= 0 && pl2.lval >= 0 && pl2.dinv >= 0 && pl2.yw >= 0 && (CD.T == 0 || pl2.dtail >= 0)) {
                // measured (profiles/r02_tuning.md section 7): a gain only while the whole shard is co-resident in the throughput
                // launch (B <= 2 x SMs: 256 instances 2 096 -> 1 940 ms over 60 rounds with a quota of 40); with several waves the
                // block scheduler already overlaps the stragglers with the next wave and the second launch only adds a tail
                // (1024 instances: 5 112 -> 5 320..5 390 ms)
                quota = h->handoff > 0 ? h->handoff : (B <= (size_t)2 * h->num_sms ? 40 : 0);
                if (quota >= h->opts.ipm_max_iter) quota = 0;
                dyn2 = (size_t)pl2.total * sizeof(double);
            }
        }
        O.handoff_k = quota;
        O.order = (h->order_user && hmode != 1) ? h->d_order : nullptr;
        const bool use_ilv = ipm && h->G > 1 && (phase == SQPQP_PHASE_FR ? h->has_ilv_fr : h->has_ilv);
        if (use_ilv) {  // G instances interleaved per CTA (ilv.cuh); flags what it cannot finish for the ADMM launch below
            CUDA_OK(launch_ilv(h, O, phase, phase == SQPQP_PHASE_FR ? h->ilv_fr : h->ilv, phase == SQPQP_PHASE_FR ? h->ilv_fr_dyn : h->ilv_dyn));
            h->launches++;
            h->last_kernel = "k_solve_ilv<" + std::to_string(h->G) + "," + std::to_string(h->ilv_nt) + "," + std::to_string(h->ilv_minb) + ">";
        } else if (ipm) {
            launch(1, O);
            if (quota > 0 && hmode == 0) {
                DevOpts O2 = O;
                O2.handoff_k = 0; O2.resume = 1;
                int t2 = pick_threads(h, phase);
                if (t2 > 512) t2 = 512;
                k_solve_cta<512, 1, 1><<<grid, t2, dyn2, st>>>(P, O2, phase, pl2);
                h->launches++;
            } else if (quota > 0 && (hmode == 1 || hmode == 3)) {
                // second stage in the same launch configuration: every instance the quota stopped, longest predicted first
                DevOpts O2 = O;
                O2.handoff_k = 0; O2.resume = 1; O2.order = nullptr;
                if (hmode == 1 && B <= 16384) {
                    k_rank_unfinished<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(P.ipm_state, P.fb_flag, h->d_order, (int)B, h->opts.ipm_rho0);
                    h->launches++;
                    O2.order = h->d_order;
                }
                launch(1, O2);
            }
            h->last_kernel = cfg == 4 ? "k_solve_cta<256,4,1>" : (cfg == 2 ? (threads <= 384 ? "k_solve_cta<384,2,1>" : "k_solve_cta<512,2,1>") : (pl.ring >= 0 ? "k_solve_cta<512,1,1>+ring" : "k_solve_cta<512,1,1>"));
            if (quota > 0) h->last_kernel += hmode == 0 ? "+handoff" : (hmode == 1 ? "+resume(ranked)" : (hmode == 3 ? "+resume" : "+stop"));
        } else {  // no factorisation available: every instance is "flagged" (non-zero) for the ADMM launch
            h->last_kernel = "k_solve_cta<..,2> (ADMM)";
            CUDA_OK(cudaMemsetAsync(P.fb_flag, 1, B * sizeof(int), st));
        }
        if (h->opts.method != 2) launch(2, O);
        h->launches--;  // counted below
    }
    h->launches++;
    return 0;
}

static int launch_solve(sqpqp_handle h, int phase) {
    int rc = begin_solve_timing(h);
    if (rc) return rc;
    rc = launch_solve_kernels(h, phase, h->stream);
    if (rc) return rc;
    return end_solve_timing(h);
}

// One SQP round of a batch whose instances are in different phases (compute_step!, sqp_trust_region.jl:370-380, per instance):
// the QP-phase launch over `act_qp` and the restoration-phase launch over `act_fr` touch disjoint instances, so they run
// CONCURRENTLY -- the restoration kernel on the high-priority side stream, forked behind everything queued on the main stream and
// joined back into it.  Measured on the 1024-instance batch (tools/gpu_phase_share.py): a round has at most a few instances in
// restoration, and their launch, run after the QP launch, kept ONE CTA busy for 9-20 ms while the GPU idled -- 58 restoration
// solves of 102 k subproblems were 8.8 % of the solve time of the whole run.
static int launch_solve_mixed(sqpqp_handle h, const int* act_qp, const int* act_fr) {
    Prob& P = h->P;
    const bool concurrent = P.has_chol && P.has_chol_fr && h->opts.method != 1 && h->handoff_mode == 0 && !(h->G > 1) &&
                            solve_team(h, SQPQP_PHASE_QP) == 1 && solve_team(h, SQPQP_PHASE_FR) == 1;
    int rc = begin_solve_timing(h);
    if (rc) return rc;
    if (concurrent) {
        CUDA_OK(cudaEventRecord(h->ev_fork, h->stream));
        CUDA_OK(cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
        P.active = act_fr;
        rc = launch_solve_kernels(h, SQPQP_PHASE_FR, h->stream2);
        if (rc) return rc;
        CUDA_OK(cudaEventRecord(h->ev_join, h->stream2));
        P.active = act_qp;
        rc = launch_solve_kernels(h, SQPQP_PHASE_QP, h->stream);
        if (rc) return rc;
        CUDA_OK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
        h->last_kernel += " || restoration phase on the side stream";
    } else {
        P.active = act_qp;
        rc = launch_solve_kernels(h, SQPQP_PHASE_QP, h->stream);
        if (rc) return rc;
        P.active = act_fr;
        rc = launch_solve_kernels(h, SQPQP_PHASE_FR, h->stream);
        if (rc) return rc;
    }
    P.active = nullptr;
    return end_solve_timing(h);
}


extern "C" int sqpqp_solve_tr(sqpqp_handle h, int32_t phase, const double* x_k, const double* delta, const double* E_override,
                              const int32_t* active, double* p, double* lambda, double* mult_x_L, double* mult_x_U,
                              double* slack, int32_t* moi_status, sqpqp_info* info) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done || !h->updated) return fail(h, SQPQP_E_STATE, "setup/update not called");
    if (phase < 0 || phase > 3 || !x_k || !delta) return fail(h, SQPQP_E_BADARG, "bad phase or null pointer");
    if (phase == SQPQP_PHASE_SOC && !E_override) return fail(h, SQPQP_E_BADARG, "SOC phase needs E_override");
    DeviceGuard g(h->device);
    Prob& P = h->P;
    const size_t B = P.batch;
    size_t bytes = B * ((size_t)3 * P.n + 2 * (size_t)P.m + P.S + 8) * sizeof(double) * 2 + B * (sizeof(sqpqp_info) + 16) + 16384;
    int rc = ensure_stage(h, bytes);
    if (rc) return rc;
    for (size_t b = 0; b < B; ++b)
        if (!(delta[b] > 0.0)) return fail(h, SQPQP_E_BADARG, "delta must be positive");
    P.xk = upload(h, x_k, B * P.n);
    P.delta = upload(h, delta, B);
    P.Eov = (phase == SQPQP_PHASE_SOC) ? upload(h, E_override, B * P.m) : nullptr;
    P.active = active ? upload(h, active, B) : nullptr;
    rc = launch_solve(h, phase);
    if (rc) return rc;
    download(h, P.o_p, p, B * P.n);
    download(h, P.o_lam, lambda, B * P.m);
    download(h, P.o_mxL, mult_x_L, B * P.n);
    download(h, P.o_mxU, mult_x_U, B * P.n);
    download(h, P.o_slack, slack, B * P.S);
    download(h, P.o_info, info, B);
    std::vector<sqpqp_info> tmp;
    if (moi_status && !info) {
        tmp.resize(B);
        download(h, P.o_info, tmp.data(), B);
    }
    rc = finish(h);
    if (rc) return rc;
    if (moi_status) {
        const sqpqp_info* src = info ? info : tmp.data();
        for (size_t b = 0; b < B; ++b) moi_status[b] = (!active || active[b]) ? src[b].moi_status : moi_status[b];
    }
    P.active = nullptr;
    P.Eov = nullptr;
    return 0;
}

// Device-pointer variant: nothing crosses PCIe, nothing blocks.  Outputs stay in the handle's
// device buffers (sqpqp_device_outputs); call sqpqp_sync before reading them.
extern "C" int sqpqp_solve_tr_device(sqpqp_handle h, int32_t phase, const double* x_k, const double* delta,
                                     const double* E_override, const int32_t* active) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done || !h->updated) return fail(h, SQPQP_E_STATE, "setup/update not called");
    if (phase < 0 || phase > 3 || !x_k || !delta) return fail(h, SQPQP_E_BADARG, "bad phase or null pointer");
    if (phase == SQPQP_PHASE_SOC && !E_override) return fail(h, SQPQP_E_BADARG, "SOC phase needs E_override");
    DeviceGuard g(h->device);
    Prob& P = h->P;
    P.xk = x_k; P.delta = delta; P.Eov = (phase == SQPQP_PHASE_SOC) ? E_override : nullptr; P.active = active;
    int rc = launch_solve(h, phase);
    P.active = nullptr;
    P.Eov = nullptr;
    return rc;
}

// Both phases of one SQP round in one call: instances with active_qp[b] != 0 solve the QP subproblem (sub_optimize!), instances
// with active_fr[b] != 0 the restoration LP (sub_optimize_FR!); the two sets must be disjoint, every other instance is skipped.
// One upload of x_k / delta, the two launches side by side (launch_solve_mixed), one download of the results.
extern "C" int sqpqp_solve_tr_mixed(sqpqp_handle h, const double* x_k, const double* delta, const int32_t* active_qp,
                                    const int32_t* active_fr, double* p, double* lambda, double* mult_x_L, double* mult_x_U,
                                    double* slack, int32_t* moi_status, sqpqp_info* info) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done || !h->updated) return fail(h, SQPQP_E_STATE, "setup/update not called");
    if (!x_k || !delta || !active_qp || !active_fr) return fail(h, SQPQP_E_BADARG, "null pointer");
    DeviceGuard g(h->device);
    Prob& P = h->P;
    const size_t B = P.batch;
    for (size_t b = 0; b < B; ++b) {
        if (active_qp[b] && active_fr[b]) return fail(h, SQPQP_E_BADARG, "an instance is in one phase per round: active_qp and active_fr overlap");
        if (!(delta[b] > 0.0)) return fail(h, SQPQP_E_BADARG, "delta must be positive");
    }
    size_t bytes = B * ((size_t)3 * P.n + 2 * (size_t)P.m + P.S + 8) * sizeof(double) * 2 + B * (sizeof(sqpqp_info) + 32) + 16384;
    int rc = ensure_stage(h, bytes);
    if (rc) return rc;
    P.xk = upload(h, x_k, B * P.n);
    P.delta = upload(h, delta, B);
    P.Eov = nullptr;
    const int* aq = upload(h, active_qp, B);
    const int* af = upload(h, active_fr, B);
    rc = launch_solve_mixed(h, aq, af);
    if (rc) return rc;
    download(h, P.o_p, p, B * P.n);
    download(h, P.o_lam, lambda, B * P.m);
    download(h, P.o_mxL, mult_x_L, B * P.n);
    download(h, P.o_mxU, mult_x_U, B * P.n);
    download(h, P.o_slack, slack, B * P.S);
    download(h, P.o_info, info, B);
    std::vector<sqpqp_info> tmp;
    if (moi_status && !info) {
        tmp.resize(B);
        download(h, P.o_info, tmp.data(), B);
    }
    rc = finish(h);
    if (rc) return rc;
    if (moi_status) {
        const sqpqp_info* src = info ? info : tmp.data();
        for (size_t b = 0; b < B; ++b) moi_status[b] = (active_qp[b] || active_fr[b]) ? src[b].moi_status : moi_status[b];
    }
    return 0;
}

// Device-pointer variant of sqpqp_solve_tr_mixed (nothing crosses PCIe, nothing blocks; the masks are device arrays and the
// caller guarantees that they are disjoint).
extern "C" int sqpqp_solve_tr_mixed_device(sqpqp_handle h, const double* x_k, const double* delta, const int32_t* active_qp,
                                           const int32_t* active_fr) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done || !h->updated) return fail(h, SQPQP_E_STATE, "setup/update not called");
    if (!x_k || !delta || !active_qp || !active_fr) return fail(h, SQPQP_E_BADARG, "null pointer");
    DeviceGuard g(h->device);
    Prob& P = h->P;
    P.xk = x_k; P.delta = delta; P.Eov = nullptr;
    return launch_solve_mixed(h, active_qp, active_fr);
}

extern "C" int sqpqp_sync(sqpqp_handle h) {
    if (!h) return SQPQP_E_BADARG;
    DeviceGuard g(h->device);
    CUDA_OK(cudaStreamSynchronize(h->stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int sqpqp_device_outputs(sqpqp_handle h, double** p, double** lambda, double** mult_x_L, double** mult_x_U,
                                    sqpqp_info** info) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done) return fail(h, SQPQP_E_STATE, "setup not called");
    if (p) *p = h->P.o_p;
    if (lambda) *lambda = h->P.o_lam;
    if (mult_x_L) *mult_x_L = h->P.o_mxL;
    if (mult_x_U) *mult_x_U = h->P.o_mxU;
    if (info) *info = h->P.o_info;
    return 0;
}

extern "C" int sqpqp_fetch_info(sqpqp_handle h, sqpqp_info* info) {
    if (!h || !info) return SQPQP_E_BADARG;
    if (!h->setup_done) return fail(h, SQPQP_E_STATE, "setup not called");
    DeviceGuard g(h->device);
    CUDA_OK(cudaMemcpyAsync(info, h->P.o_info, h->P.batch * sizeof(sqpqp_info), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---- merit / KT --------------------------------------------------------------------------
extern "C" int sqpqp_merit(sqpqp_handle h, const double* x, const double* p, const double* E_trial, const double* f_trial,
                           const double* mu, const int32_t* fr, double* viol0, double* viol_trial, double* phi_trial,
                           double* q0, double* qk) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done || !h->updated) return fail(h, SQPQP_E_STATE, "setup/update not called");
    if (!x || !p || !mu) return fail(h, SQPQP_E_BADARG, "null pointer");
    // E_trial / f_trial == NULL: the values sqpqp_acopf_eval_trial left on the device
    if ((!E_trial || !f_trial) && !h->trial_valid) return fail(h, SQPQP_E_BADARG, "null trial values without a preceding sqpqp_acopf_eval_trial");
    DeviceGuard g(h->device);
    Prob& P = h->P;
    const size_t B = P.batch;
    int rc = ensure_stage(h, B * ((size_t)2 * P.n + P.m + 16) * sizeof(double) + 16384);
    if (rc) return rc;
    MeritArgs A;
    A.x = upload(h, x, B * P.n);
    A.p = upload(h, p, B * P.n);
    A.Etrial = E_trial ? upload(h, E_trial, B * P.m) : h->d_Etrial;
    A.ftrial = f_trial ? upload(h, f_trial, B) : h->d_ftrial;
    A.mu = upload(h, mu, B);
    A.fr = fr ? upload(h, fr, B) : nullptr;
    double* out;  // 5 x B results in the (zeroed) N_TMP workspace of instance 0.. (batch*Ne >= 5*batch only if Ne>=5)
    size_t off = h->stage_off;
    h->stage_off = align256(off + 5 * B * sizeof(double));
    out = (double*)(h->dstage + off);
    A.viol0 = out; A.violt = out + B; A.phit = out + 2 * B; A.q0 = out + 3 * B; A.qk = out + 4 * B;
    P.active = nullptr;
    k_merit<<<(int)(B < 65535 ? B : 65535), 128, 0, h->stream>>>(P, A);
    h->launches++;
    download(h, A.viol0, viol0, B);
    download(h, A.violt, viol_trial, B);
    download(h, A.phit, phi_trial, B);
    download(h, A.q0, q0, B);
    download(h, A.qk, qk, B);
    return finish(h);
}

extern "C" int sqpqp_linesearch_terms(sqpqp_handle h, const double* x, const double* p, const double* alpha, const double* E_trial,
                                      const double* mu_rows, const double* lambda, double* out8) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done || !h->updated) return fail(h, SQPQP_E_STATE, "setup/update not called");
    if (!x || !p || !alpha || !E_trial || !mu_rows || !lambda || !out8) return fail(h, SQPQP_E_BADARG, "null pointer");
    DeviceGuard g(h->device);
    Prob& P = h->P;
    const size_t B = P.batch;
    int rc = ensure_stage(h, B * ((size_t)2 * P.n + 3 * (size_t)P.m + 32) * sizeof(double) * 2 + 16384);
    if (rc) return rc;
    LsArgs A;
    A.x = upload(h, x, B * P.n);
    A.p = upload(h, p, B * P.n);
    A.alpha = upload(h, alpha, B);
    A.Etrial = upload(h, E_trial, B * P.m);
    A.mu = upload(h, mu_rows, B * P.m);
    A.lam = upload(h, lambda, B * P.m);
    size_t off = h->stage_off;
    h->stage_off = align256(off + 8 * B * sizeof(double));
    A.out = (double*)(h->dstage + off);
    P.active = nullptr;
    k_linesearch<<<(int)(B < 65535 ? B : 65535), 128, 0, h->stream>>>(P, A);
    h->launches++;
    download(h, A.out, out8, 8 * B);
    return finish(h);
}

extern "C" int sqpqp_kt_residuals(sqpqp_handle h, const double* lambda, const double* mult_x_U, const double* mult_x_L, double* kt) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done || !h->updated) return fail(h, SQPQP_E_STATE, "setup/update not called");
    if (!lambda || !mult_x_U || !mult_x_L || !kt) return fail(h, SQPQP_E_BADARG, "null pointer");
    DeviceGuard g(h->device);
    Prob& P = h->P;
    const size_t B = P.batch;
    int rc = ensure_stage(h, B * ((size_t)2 * P.n + P.m + 8) * sizeof(double) * 2 + 16384);
    if (rc) return rc;
    KtArgs A;
    A.lam = upload(h, lambda, B * P.m);
    A.mxU = upload(h, mult_x_U, B * P.n);
    A.mxL = upload(h, mult_x_L, B * P.n);
    size_t off = h->stage_off;
    h->stage_off = align256(off + B * sizeof(double));
    A.kt = (double*)(h->dstage + off);
    P.active = nullptr;
    k_kt<<<(int)(B < 65535 ? B : 65535), 128, 0, h->stream>>>(P, A);
    h->launches++;
    download(h, A.kt, kt, B);
    return finish(h);
}

// which: 0 y = J x (m <- n), 1 y = J' x (n <- m), 2 y = H x (n <- n); device pointers, [batch][len]
static int launch_spmv(sqpqp_handle h, int which, const double* x, double* y) {
    Prob& P = h->P;
    const SpmvPlan& S = which == 0 ? h->planJ : (which == 1 ? h->planT : h->planH);
    const double* vals = which == 0 ? P.Jv : (which == 1 ? P.Tv : P.Hv);
    if (S.nblocks <= 0 || S.nrows <= 0) return 0;
    // one CTA per (row block, pair of instances): many short-lived CTAs overlap their load and reduce phases better than
    // persistent ones (measured: 8 resident CTAs per SM walking the instances reach 1.9 TB/s, this grid 2.5-3.1 TB/s)
    int gy = (P.batch + SPMV_INST - 1) / SPMV_INST;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    k_spmv_stream<<<dim3(S.nblocks, gy), SPMV_THREADS, 0, h->stream>>>(S, vals, x, y, which == 1 ? P.m : P.n, P.batch);
    h->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int sqpqp_jac_times(sqpqp_handle h, const double* p, double* out) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done || !h->updated) return fail(h, SQPQP_E_STATE, "setup/update not called");
    if (!p || !out) return fail(h, SQPQP_E_BADARG, "null pointer");
    DeviceGuard g(h->device);
    Prob& P = h->P;
    const size_t B = P.batch;
    int rc = ensure_stage(h, B * ((size_t)P.n + 2 * (size_t)P.m + 8) * sizeof(double) + 16384);
    if (rc) return rc;
    const double* dp = upload(h, p, B * P.n);
    size_t off = h->stage_off;
    h->stage_off = align256(off + B * P.m * sizeof(double));
    double* dout = (double*)(h->dstage + off);
    rc = launch_spmv(h, 0, dp, dout);
    if (rc) return rc;
    download(h, dout, out, B * P.m);
    return finish(h);
}

extern "C" int sqpqp_spmv_device(sqpqp_handle h, int32_t which, const double* x_dev, double* y_dev) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done || !h->updated) return fail(h, SQPQP_E_STATE, "setup/update not called");
    if (which < 0 || which > 2 || !x_dev || !y_dev) return fail(h, SQPQP_E_BADARG, "bad selector or null pointer");
    DeviceGuard g(h->device);
    if (h->timing_pending) sqpqp_last_solve_ms(h);
    h->ev0 = h->sev[0]; h->ev1 = h->sev[1];
    CUDA_OK(cudaEventRecord(h->ev0, h->stream));
    int rc = launch_spmv(h, which, x_dev, y_dev);
    if (rc) return rc;
    CUDA_OK(cudaEventRecord(h->ev1, h->stream));
    h->timing_pending = true;
    h->timing_is_solve = false;
    return 0;
}

extern "C" int sqpqp_spmv(sqpqp_handle h, int32_t which, const double* x, double* y) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done || !h->updated) return fail(h, SQPQP_E_STATE, "setup/update not called");
    if (which < 0 || which > 2 || !x || !y) return fail(h, SQPQP_E_BADARG, "bad selector or null pointer");
    DeviceGuard g(h->device);
    Prob& P = h->P;
    const size_t B = P.batch, nx = which == 1 ? P.m : P.n, ny = which == 0 ? P.m : P.n;
    int rc = ensure_stage(h, B * (nx + ny + 8) * sizeof(double) * 2 + 16384);
    if (rc) return rc;
    const double* dx = upload(h, x, B * nx);
    size_t off = h->stage_off;
    h->stage_off = align256(off + B * ny * sizeof(double));
    double* dy = (double*)(h->dstage + off);
    rc = launch_spmv(h, which, dx, dy);
    if (rc) return rc;
    download(h, dy, y, B * ny);
    return finish(h);
}

// ---- read-back of the device matrices (parity tests) ----------------------------------------
extern "C" int sqpqp_get_csr(sqpqp_handle h, int32_t which, int32_t b, int64_t* nnz, int32_t* row_ptr, int32_t* col_idx, double* values) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done) return fail(h, SQPQP_E_STATE, "setup not called");
    Prob& P = h->P;
    if (which < 0 || which > 2 || b < 0 || b >= P.batch) return fail(h, SQPQP_E_BADARG, "bad selector");
    DeviceGuard g(h->device);
    CUDA_OK(cudaStreamSynchronize(h->stream));
    int nrows = which == 0 ? P.m : P.n;
    const int* rb = which == 0 ? P.J_rb : (which == 1 ? P.T_rb : P.H_rb);
    const int* col = which == 0 ? P.J_col : (which == 1 ? P.T_col : P.H_col);
    const double* val = which == 0 ? P.Jv + (size_t)b * P.nnzJ : (which == 1 ? P.Tv + (size_t)b * P.nnzT : P.Hv + (size_t)b * P.nnzH);
    std::vector<int> rp(nrows + 1);
    CUDA_OK(cudaMemcpy(rp.data(), rb, (nrows + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    if (which == 0) {
        // strip the slack columns: row i keeps [rb[i], re_n[i])
        std::vector<int> ren(nrows);
        if (nrows) CUDA_OK(cudaMemcpy(ren.data(), P.J_re_n, nrows * sizeof(int), cudaMemcpyDeviceToHost));
        std::vector<int> c(P.nnzJ);
        std::vector<double> v(P.nnzJ);
        if (P.nnzJ) {
            CUDA_OK(cudaMemcpy(c.data(), col, P.nnzJ * sizeof(int), cudaMemcpyDeviceToHost));
            CUDA_OK(cudaMemcpy(v.data(), val, P.nnzJ * sizeof(double), cudaMemcpyDeviceToHost));
        }
        int64_t cnt = 0;
        for (int i = 0; i < nrows; ++i) {
            if (row_ptr) row_ptr[i] = (int32_t)cnt;
            for (int k = rp[i]; k < ren[i]; ++k, ++cnt) {
                if (col_idx) col_idx[cnt] = c[k];
                if (values) values[cnt] = v[k];
            }
        }
        if (row_ptr) row_ptr[nrows] = (int32_t)cnt;
        if (nnz) *nnz = cnt;
        return 0;
    }
    int64_t cnt = rp[nrows];
    if (nnz) *nnz = cnt;
    if (row_ptr) memcpy(row_ptr, rp.data(), (nrows + 1) * sizeof(int));
    if (col_idx && cnt) CUDA_OK(cudaMemcpy(col_idx, col, cnt * sizeof(int), cudaMemcpyDeviceToHost));
    if (values && cnt) CUDA_OK(cudaMemcpy(values, val, cnt * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

// ---- generic QP lane ------------------------------------------------------------------------
extern "C" int sqpqp_qp_setup(sqpqp_handle h, int32_t nv, int32_t nc, int64_t nnz_p, const int64_t* p_row, const int64_t* p_col,
                              int64_t nnz_a, const int64_t* a_row, const int64_t* a_col) {
    if (!h) return SQPQP_E_BADARG;
    if (nv < 1 || nc < 0) return fail(h, SQPQP_E_BADARG, "bad sizes");
    // bounds are supplied per solve; setup only needs finiteness for slack columns, and with
    // m_lin = nc there are none.
    std::vector<double> lo((size_t)(nv > nc ? nv : nc), -INFINITY), hi((size_t)(nv > nc ? nv : nc), INFINITY);
    int rc = sqpqp_setup_nlp(h, 1, nv, nc, nc, nnz_a, a_row, a_col, nnz_p, p_row, p_col, lo.data(), hi.data(), lo.data(), hi.data(), 0);
    if (rc) return rc;
    h->generic = true;
    return 0;
}

extern "C" int sqpqp_qp_solve(sqpqp_handle h, const double* p_val, const double* q, const double* a_val, const double* rl,
                              const double* ru, const double* cl, const double* cu, double* x, double* row_dual,
                              double* col_dual, int32_t* moi_status, sqpqp_info* info) {
    if (!h) return SQPQP_E_BADARG;
    if (!h->setup_done || !h->generic) return fail(h, SQPQP_E_STATE, "qp_setup not called");
    Prob& P = h->P;
    if (!q || !cl || !cu || (P.m && (!rl || !ru))) return fail(h, SQPQP_E_BADARG, "null pointer");
    DeviceGuard g(h->device);
    // bounds
    int rc = ensure_stage(h, ((size_t)2 * P.n + 2 * (size_t)P.m) * sizeof(double) + 8192);
    if (rc) return rc;
    CUDA_OK(cudaMemcpyAsync(h->d_xL, upload(h, cl, (size_t)P.n), P.n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    CUDA_OK(cudaMemcpyAsync(h->d_xU, upload(h, cu, (size_t)P.n), P.n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (P.m) {
        CUDA_OK(cudaMemcpyAsync(h->d_gL, upload(h, rl, (size_t)P.m), P.m * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CUDA_OK(cudaMemcpyAsync(h->d_gU, upload(h, ru, (size_t)P.m), P.m * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    }
    CUDA_OK(cudaStreamSynchronize(h->stream));
    std::vector<double> zeros((size_t)(P.n > P.m ? P.n : P.m), 0.0);
    rc = sqpqp_update_nlp(h, a_val, p_val, q, zeros.data());
    if (rc) return rc;
    double delta = INFINITY;
    std::vector<double> mxL(P.n), mxU(P.n);
    int32_t st = 0;
    rc = sqpqp_solve_tr(h, SQPQP_PHASE_QP, zeros.data(), &delta, nullptr, nullptr, x, row_dual, mxL.data(), mxU.data(), nullptr, &st, info);
    if (rc) return rc;
    if (col_dual)
        for (int j = 0; j < P.n; ++j) col_dual[j] = mxL[j] + mxU[j];
    if (moi_status) *moi_status = st;
    return 0;
}
