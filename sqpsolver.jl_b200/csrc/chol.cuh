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
//   sparse levels (columns 0..n0-1): two barrier phases per elimination-tree level (diagonal, then
//       sub-diagonal entries), each a flat task list with one coalesced descriptor per task; a
//       sub-warp per task, its multiply-subtract pairs gathered four chains at a time.  The
//       dependent chain of a phase is descriptor -> pair list -> L values.
//   dense tail (last T columns = the chain at the top of the tree, where L is nearly dense):
//       its Schur complement is gathered in ONE parallel phase into a packed T x T lower triangle
//       in shared memory and factorised there by a left-looking dense Cholesky (one shared-memory
//       barrier per column, no global traffic); the triangular solves on it run in registers of
//       one warp with shuffle broadcasts.
// Deterministic (no atomics); a non-positive pivot is reported, never hidden.
#pragma once
#include "team.cuh"

// lanes (power of two <= 32) that cooperate on one task of a phase with `count` independent
// tasks: wide when the phase is narrow, 1 when it has at least as many tasks as the team has threads.
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

// sum_{q = q0, q0 + step, ... < qe} A[ab[q].x] * B[ab[q].y] with four independent gather chains in
// flight (index -> value is a dependent L2 round trip; a plain loop would pay it once per pair).
__device__ __forceinline__ double gather_dot(const int2* __restrict__ ab, int q, const int qe, const int step,
                                             const double* A, const double* B) {
    double acc0 = 0.0, acc1 = 0.0;
    for (; q + 3 * step < qe; q += 4 * step) {
        const int2 p0 = ab[q], p1 = ab[q + step], p2 = ab[q + 2 * step], p3 = ab[q + 3 * step];
        const double a0 = A[p0.x], b0 = B[p0.y], a1 = A[p1.x], b1 = B[p1.y];
        const double a2 = A[p2.x], b2 = B[p2.y], a3 = A[p3.x], b3 = B[p3.y];
        acc0 = fma(a0, b0, acc0); acc1 = fma(a1, b1, acc1);
        acc0 = fma(a2, b2, acc0); acc1 = fma(a3, b3, acc1);
    }
    if (q < qe) {
        const bool h1 = q + step < qe, h2 = q + 2 * step < qe;
        const int2 p0 = ab[q], p1 = h1 ? ab[q + step] : p0, p2 = h2 ? ab[q + 2 * step] : p0;
        const double a0 = A[p0.x], b0 = B[p0.y], a1 = A[p1.x], b1 = B[p1.y], a2 = A[p2.x], b2 = B[p2.y];
        acc0 = fma(a0, b0, acc0);
        if (h1) acc1 = fma(a1, b1, acc1);
        if (h2) acc0 = fma(a2, b2, acc0);
    }
    return acc0 + acc1;
}
// sum_{p = p0, p0 + step, ... < pe} V[p] * Y[idx[p]]  (a column of L against the solve scratch)
__device__ __forceinline__ double column_dot(const double* __restrict__ V, const int* __restrict__ idx, int p, const int pe,
                                             const int step, const double* Y) {
    double acc0 = 0.0, acc1 = 0.0;
    for (; p + 3 * step < pe; p += 4 * step) {
        const int i0 = idx[p], i1 = idx[p + step], i2 = idx[p + 2 * step], i3 = idx[p + 3 * step];
        const double v0 = V[p], v1 = V[p + step], v2 = V[p + 2 * step], v3 = V[p + 3 * step];
        acc0 = fma(v0, Y[i0], acc0); acc1 = fma(v1, Y[i1], acc1);
        acc0 = fma(v2, Y[i2], acc0); acc1 = fma(v3, Y[i3], acc1);
    }
    for (; p < pe; p += step) acc0 = fma(V[p], Y[idx[p]], acc0);
    return acc0 + acc1;
}

// Two pair lists walked together, two pairs of each per trip: eight independent gathers in flight.  A thread that owns
// two slots of a phase pays the index -> value round trips once for both.
__device__ __forceinline__ void gather_dot2(const int2* __restrict__ ab, int qa, const int ea, const int sa, int qb, const int eb,
                                            const int sb, const double* A, const double* Ba, const double* Bb, double& ra, double& rb) {
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
    while (qa < ea || qb < eb) {
        const bool a_1 = qa < ea, a_2 = qa + sa < ea, b_1 = qb < eb, b_2 = qb + sb < eb;
        const int2 pa1 = ab[a_1 ? qa : 0], pa2 = ab[a_2 ? qa + sa : 0], pb1 = ab[b_1 ? qb : 0], pb2 = ab[b_2 ? qb + sb : 0];
        const double xa1 = A[pa1.x], ya1 = Ba[pa1.y], xa2 = A[pa2.x], ya2 = Ba[pa2.y];
        const double xb1 = A[pb1.x], yb1 = Bb[pb1.y], xb2 = A[pb2.x], yb2 = Bb[pb2.y];
        if (a_1) a0 = fma(xa1, ya1, a0);
        if (a_2) a1 = fma(xa2, ya2, a1);
        if (b_1) b0 = fma(xb1, yb1, b0);
        if (b_2) b1 = fma(xb2, yb2, b1);
        qa += 2 * sa;
        qb += 2 * sb;
    }
    ra = a0 + a1;
    rb = b0 + b1;
}


// ---- shared-memory ring of index-program chunks (resident CTA team) ----------------------------------------------------
// One CTA owns an SM and keeps L, the dense tail, the solve scratch and the gathered vectors in shared memory; what is
// left on the critical path of a barrier phase of the slot-list code above is the L2 round trip of its index program
// (slot -> pair list, ~3 000 cycles per phase measured).  The program is static, so symbolic.hpp lays it out as
// self-contained chunk images and thread 0 streams them into RING_S shared-memory stages with bulk asynchronous copies
// (cp.async.bulk, completion on an mbarrier) RING_S - 1 chunks ahead of the consumers.  A chunk is one strip of slots of one
// barrier phase; every chunk ends with the __syncthreads the phase needs anyway, which is also what frees its stage.
extern __shared__ double ring_sm[];  // the CTA's dynamic shared memory (same base as dsm in k_solve_cta): shared-space loads
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ring_init(Ring& R, const CholDev& C, const CholWork& W, int stage0_dbl, unsigned long long* bars) {
    R.prog = C.rprog; R.stage0w = 2 * stage0_dbl; R.stage_words = C.ring_stage_words; R.bar0 = smem_u32(bars);
    R.phase_bits = 0; R.head = 0; R.inflight = 0; R.seg = -1; R.on = true;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < RING_S; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(R.bar0 + 8 * s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ring_sm[W.oL + C.ring_nL] = 0.0;  // the zero entry short pair lists are padded with
    }
    __syncthreads();
}
// thread 0 only: bulk copy of one chunk image into a stage, completion (byte count) on the stage's barrier
__device__ __forceinline__ void ring_issue(const Ring& R, int stage, int word_off, int bytes) {
    const unsigned bar = R.bar0 + 8 * stage;
    const unsigned dst = smem_u32(reinterpret_cast<const int*>(ring_sm) + R.stage0w + stage * R.stage_words);
    const int* src = R.prog + word_off;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void ring_wait_stage(Ring& R, int stage) {
    const unsigned bar = R.bar0 + 8 * stage, par = (R.phase_bits >> stage) & 1u;
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(par) : "memory");
    }
    R.phase_bits ^= 1u << stage;
}
// Make `seg` the segment whose first chunks are in flight.  Called by all threads, after a barrier that follows the last
// read of any stage.  A segment primed earlier but not wanted is drained first: a bulk copy cannot be cancelled.
__device__ __forceinline__ void ring_prime(Ring& R, const CholDev& C, int seg) {
    if (R.seg == seg && R.head == 0 && R.inflight > 0) return;
    if (R.inflight > 0) {  // (uniform) every thread must have seen these phases complete before their barriers are re-armed
        for (int i = 0; i < R.inflight; ++i) ring_wait_stage(R, (R.head + i) % RING_S);
        __syncthreads();
    }
    const int cnt = C.rseg_n[seg] < RING_S ? C.rseg_n[seg] : RING_S;
    if (threadIdx.x == 0)
        for (int s = 0; s < cnt; ++s) ring_issue(R, s, C.rseg_off[seg][s], C.rseg_bytes[seg][s]);
    R.seg = seg; R.head = 0; R.inflight = cnt;
}
__device__ __forceinline__ void ring_drain(Ring& R) {
    for (int i = 0; i < R.inflight; ++i) ring_wait_stage(R, (R.head + i) % RING_S);
    R.inflight = 0; R.head = 0; R.seg = -1;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < RING_S; ++s) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(R.bar0 + 8 * s));
    }
    R.on = false;
}

struct RingArgs {
    const double* Pv;   // scaled P values or null
    const double* dg;   // diagonal term per ORIGINAL column
    double shift;       // inertia shift added to every diagonal entry
    const double* Jv;   // scaled J values (assembly)
};
// One segment of the ring program (symbolic.hpp: kinds 0..6).  Returns 1.0 if a diagonal pivot was not positive (per thread;
// the caller reduces).  `next_seg` >= 0: its first chunks are requested as soon as the last stage of this segment is free.
// Factor / sweep chunks read L, yw, dinv, D and the chunk itself through the shared window (32-bit addresses, no
// predicates: short pair lists multiply the zero entry); assembly chunks gather wJ and Jv from global memory.
// The ring state and the work pointers are copied into locals for the duration of the segment (both structs live in the
// caller's frame; behind a reference every asm with a memory clobber would force their fields to be re-read).
__device__ __forceinline__ double ring_run_segment(Ring& Rref, const CholDev& C, const CholWork& Wref, int seg, int next_seg,
                                                   const RingArgs& aref, Prof& pf) {
    Ring R = Rref;
    const CholWork W = Wref;
    const RingArgs a = aref;
    ring_prime(R, C, seg);
    double bad = 0.0;
    int prev_asm = 0;
    const int* smw = reinterpret_cast<const int*>(ring_sm);
    const int oL = W.oL, oyw = W.oyw, odinv = W.odinv, oD = W.oD;
    for (;;) {
        const int stage = R.head % RING_S;
        const long long c0 = SQPQP_CLK();
        ring_wait_stage(R, stage);
        const long long c1 = SQPQP_CLK();
        const int* img = smw + R.stage0w + stage * R.stage_words;
        const int npad = img[0], kmax = img[1], is_asm = img[2], last = img[3], nxo = img[4], nxb = img[5];
        if (prev_asm && !is_asm) pf.lap(PS_ASSEMBLE_SLOTS);
        prev_asm = is_asm;
        const int2* slots = reinterpret_cast<const int2*>(img + 8);
        if (!is_asm) {
            for (int s = threadIdx.x; s < npad; s += blockDim.x) {  // whole warps: npad and blockDim are multiples of 32
                const int2 sl = slots[s];
                const int* pw = img + 8 + 2 * npad + s;
                const int tgt = sl.x & 0xffff, lgv = (sl.x >> 16) & 7, kind = (sl.x >> 28) & 7;
                const bool leader = (sl.x >> 19) & 1;
                const int oB = kind >= 3 ? oyw : oL;
                // operands of the finalisation that do not depend on the sum: issued before the gathers
                double f0 = 0.0, f1 = 1.0;
                if (leader) {
                    if (kind <= 1) { if ((sl.x >> 20) & 1) f0 = ring_sm[oL + tgt]; if (kind == 1) f1 = ring_sm[odinv + sl.y]; }
                    else if (kind == 2) f0 = ring_sm[oD + sl.y];
                    else { f0 = ring_sm[oyw + tgt]; if (kind == 3) f1 = ring_sm[odinv + tgt]; }
                }
                double acc0 = 0.0, acc1 = 0.0;
                for (int k = 0; k < kmax; k += 2) {
                    const unsigned p0 = (unsigned)pw[k * npad], p1 = (unsigned)pw[(k + 1) * npad];
                    acc0 = fma(ring_sm[oL + (p0 & 0xffffu)], ring_sm[oB + (p0 >> 16)], acc0);
                    acc1 = fma(ring_sm[oL + (p1 & 0xffffu)], ring_sm[oB + (p1 >> 16)], acc1);
                }
                double acc = acc0 + acc1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {  // lane groups of mixed (power of two, aligned) sizes share the butterfly
                    const double t = __shfl_xor_sync(0xffffffffu, acc, o);
                    if (o < (1 << lgv)) acc += t;
                }
                if (leader) {
                    double v = f0 - acc;
                    if (kind == 0) {
                        if (!(v > 0.0)) { bad = 1.0; v = 1.0; }
                        const double inv = rsqrt(v);
                        ring_sm[oL + tgt] = v * inv;
                        ring_sm[odinv + sl.y] = inv;
                    } else if (kind == 1) ring_sm[oL + tgt] = v * f1;
                    else if (kind == 2) ring_sm[oD + sl.y] = v;
                    else ring_sm[oyw + tgt] = v * f1;  // kind 3: scaled by the pivot; kind 4: f1 = 1
                }
            }
        } else {
            for (int s = threadIdx.x; s < npad; s += blockDim.x) {
                const int2 sl = slots[s];
                const int* pw = img + 8 + 2 * npad + s;
                const int tgt = sl.x & 0xffff, lgv = (sl.x >> 16) & 7, ks = (sl.x >> 21) & 0x7f, kind = (sl.x >> 28) & 7;
                const bool leader = (sl.x >> 19) & 1;
                double f0 = 0.0;
                if (leader) {
                    const int h = (sl.y & 0xffff) - 1, d = (int)((unsigned)sl.y >> 16) - 1;
                    if (a.Pv && h >= 0) f0 = a.Pv[h];
                    if (d >= 0) f0 += a.dg[d] + a.shift;
                }
                double acc0 = 0.0, acc1 = 0.0;
                for (int k = 0; k < kmax; k += 2) {
                    const unsigned p0 = (unsigned)pw[k * npad], p1 = (unsigned)pw[(k + 1) * npad];
                    const bool o0 = k < ks, o1 = k + 1 < ks;
                    const double a0 = W.wJ[o0 ? (p0 & 0xffffu) : 0u], b0 = a.Jv[o0 ? (p0 >> 16) : 0u];
                    const double a1 = W.wJ[o1 ? (p1 & 0xffffu) : 0u], b1 = a.Jv[o1 ? (p1 >> 16) : 0u];
                    if (o0) acc0 = fma(a0, b0, acc0);
                    if (o1) acc1 = fma(a1, b1, acc1);
                }
                double acc = acc0 + acc1;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double t = __shfl_xor_sync(0xffffffffu, acc, o);
                    if (o < (1 << lgv)) acc += t;
                }
                if (leader) ring_sm[(kind == 5 ? oL : oD) + tgt] = f0 + acc;
            }
        }
        const long long c2 = SQPQP_CLK();
        __syncthreads();  // the phase barrier; every thread is done with this stage
        const long long c3 = SQPQP_CLK();
        pf.add(PS_RING_WAIT, c1 - c0); pf.add(PS_RING_WORK, c2 - c1); pf.add(PS_RING_BAR, c3 - c2); pf.add(PS_RING_CHUNKS, 1);
        pf.add(20 + is_asm, c2 - c1);  // (profile builds) work cycles: factor / sweep chunks, assembly chunks
        if (nxo >= 0) { if (threadIdx.x == 0) ring_issue(R, stage, nxo, nxb); }
        else --R.inflight;
        ++R.head;
        if (last) break;
    }
    R.seg = -1; R.head = 0;  // (inflight is 0 here: the chunks without a successor each gave one back)
    if (next_seg >= 0) ring_prime(R, C, next_seg);
    Rref = R;
    return bad;
}

// L <- lower triangle of K in the permuted order, sourced entries only (pure fill entries have
// K_e = 0 and are never read: their factor task carries has_K = 0).  Pv may be null (no P); dg[j] + shift
// is added to the diagonal entry of ORIGINAL column j.  Also clears the dense tail.
template <class Team>
__device__ void chol_assemble(Team& T, const CholDev& C, const CholWork& W, const double* __restrict__ Pv,
                              const double* __restrict__ dg, const double shift, const double* __restrict__ w,
                              const double* __restrict__ Jv, Prof& pf) {
    double* L = W.L;
    double* wJ = W.wJ;
    // four slots per trip, all loads before the first store: the arrays may alias as far as the compiler knows, so a
    // plain loop would serialise the index -> weight round trips of consecutive iterations
    for (int a0 = T.tid(); a0 < C.nslotJ; a0 += 4 * T.size()) {
        int r[4];
        double jv[4], wr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int a = a0 + u * T.size();
            const bool ok = a < C.nslotJ;
            r[u] = ok ? C.jrow[a] : -1;
            jv[u] = ok ? Jv[a] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) wr[u] = r[u] >= 0 ? w[r[u]] : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int a = a0 + u * T.size();
            if (a < C.nslotJ) wJ[a] = wr[u] * jv[u];
        }
    }
    if (C.T > 0) {  // packed tail, padded to a multiple of the panel width with identity rows (4: dense_factor, 32: dense_factor_grid)
        const int Tp = C.Tpad;
        for (int i = T.tid(); i < Tp * (Tp + 1) / 2; i += T.size()) W.D[i] = 0.0;
        T.sync();
        for (int i = C.T + T.tid(); i < Tp; i += T.size()) W.D[i * (i + 1) / 2 + i] = 1.0;
    }
    T.sync();
    pf.lap(PS_ASSEMBLE);
    // one slot per thread and round: a lane's share of the terms of one sourced entry (symbolic.hpp 4b), lane groups of
    // mixed power-of-two sizes reduced by one shared butterfly, the leader adds P, the diagonal term and the shift
    for (int r0 = 0; r0 < C.n_aslot; r0 += 2 * T.size()) {  // two slots per thread and trip (gather_dot2)
        const int sA = r0 + T.tid(), sB = sA + T.size();
        const bool onA = sA < C.n_aslot, onB = sB < C.n_aslot;
        const int4 slA = C.aslot[onA ? sA : 0], slB = C.aslot[onB ? sB : 0];
        const int LnA = 1 << ((slA.x >> 26) & 7), LnB = 1 << ((slB.x >> 26) & 7);
        const bool ldA = onA && ((slA.x >> 29) & 1), ldB = onB && ((slB.x >> 29) & 1);
        double vA = 0.0, vB = 0.0;
        if (ldA) {
            const int d = C.aslot_d[sA];
            if (Pv && slA.w >= 0) vA = Pv[slA.w];
            if (d >= 0) vA += dg[d] + shift;
        }
        if (ldB) {
            const int d = C.aslot_d[sB];
            if (Pv && slB.w >= 0) vB = Pv[slB.w];
            if (d >= 0) vB += dg[d] + shift;
        }
        double accA, accB;
        gather_dot2(C.as_ab, onA ? slA.y : 0, onA ? slA.z : 0, LnA, onB ? slB.y : 0, onB ? slB.z : 0, LnB, wJ, Jv, Jv, accA, accB);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double tA = __shfl_xor_sync(0xffffffffu, accA, o), tB = __shfl_xor_sync(0xffffffffu, accB, o);
            if (o < LnA) accA += tA;
            if (o < LnB) accB += tB;
        }
        if (ldA) L[slA.x & 0x3ffffff] = vA + accA;
        if (ldB) L[slB.x & 0x3ffffff] = vB + accB;
    }
    T.sync();
    pf.lap(PS_ASSEMBLE_SLOTS);
}

// ---- dense tail, CTA teams only -------------------------------------------------------------
// packed row-major lower triangle: (r, c), c <= r, at r (r + 1) / 2 + c
__device__ __forceinline__ int tri(int r) { return (r * (r + 1)) >> 1; }

// Right-looking dense Cholesky of the packed D in panels of 4 columns.  Tp = T rounded up to a multiple
// of 4 (the padding rows are identity: chol_assemble sets them).  Per panel:
//   A. every thread reads the 4 x 4 diagonal block (10 broadcast loads) and factorises it in registers --
//      redundantly, which costs no barrier; thread r then forward-substitutes row j0 + 4 + r of the panel
//      (4 values) against it, thread 0 stores the block's factor and the inverse pivots;
//   B. after one barrier the trailing matrix gets its rank-4 update, a warp per row, a lane per column.
// Two barriers per 4 columns, no reductions (the per-column variant needs a barrier, a shuffle
// reduction and a dependent rsqrt for every column).  Returns 1.0 if a pivot was not positive
// (uniform: every thread factorises the same diagonal blocks).
__device__ inline double dense_factor(double* __restrict__ D, double* __restrict__ dinvT, int Tn) {
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, wp = tid >> 5, nw = nth >> 5;
    const int Tp = (Tn + 3) & ~3;
    double bad = 0.0;
    for (int j0 = 0; j0 < Tp; j0 += 4) {
        const double* d0 = D + tri(j0) + j0;
        const double* d1 = D + tri(j0 + 1) + j0;
        const double* d2 = D + tri(j0 + 2) + j0;
        const double* d3 = D + tri(j0 + 3) + j0;
        double a00 = d0[0], a10 = d1[0], a11 = d1[1], a20 = d2[0], a21 = d2[1], a22 = d2[2];
        double a30 = d3[0], a31 = d3[1], a32 = d3[2], a33 = d3[3];
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
        for (int i = j0 + 4 + tid; i < Tp; i += nth) {
            double* r = D + tri(i) + j0;
            const double x0 = r[0] * i0;
            const double x1 = fma(-x0, l10, r[1]) * i1;
            const double x2 = fma(-x1, l21, fma(-x0, l20, r[2])) * i2;
            const double x3 = fma(-x2, l32, fma(-x1, l31, fma(-x0, l30, r[3]))) * i3;
            r[0] = x0; r[1] = x1; r[2] = x2; r[3] = x3;
        }
        __syncthreads();  // every thread has read the diagonal block; the panel rows are written
        if (tid == 0) {
            double* w1 = D + tri(j0 + 1) + j0;
            double* w2 = D + tri(j0 + 2) + j0;
            double* w3 = D + tri(j0 + 3) + j0;
            w1[0] = l10; w2[0] = l20; w2[1] = l21; w3[0] = l30; w3[1] = l31; w3[2] = l32;
            if (j0 < Tn) dinvT[j0] = i0;
            if (j0 + 1 < Tn) dinvT[j0 + 1] = i1;
            if (j0 + 2 < Tn) dinvT[j0 + 2] = i2;
            if (j0 + 3 < Tn) dinvT[j0 + 3] = i3;
        }
        for (int i = j0 + 4 + wp; i < Tp; i += nw) {  // rank-4 update: a warp per row, a lane per column
            double* __restrict__ row = D + tri(i);
            const double xi0 = row[j0], xi1 = row[j0 + 1], xi2 = row[j0 + 2], xi3 = row[j0 + 3];
            for (int k = j0 + 4 + lane; k <= i; k += 32) {
                const double* xk = D + tri(k) + j0;
                const double s = fma(xi3, xk[3], fma(xi2, xk[2], fma(xi1, xk[1], xi0 * xk[0])));
                row[k] -= s;
            }
        }
        __syncthreads();
    }
    return bad;
}

// y <- L_T^{-1} y, then y <- L_T^{-T} y on the tail part of the solve scratch; one warp, the
// right-hand side in registers (lane owns rows lane, lane+32, ...), pivots broadcast by shuffle:
// the dependent chain of a step is shuffle -> multiply -> fma.  Tn <= 128.
__device__ inline void dense_solve_warp(const double* __restrict__ D, const double* __restrict__ dinvT, double* yt, int Tn) {
    const int lane = threadIdx.x & 31;
    double t[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) t[s] = (lane + 32 * s < Tn) ? yt[lane + 32 * s] : 0.0;
    // forward, column oriented: y_j = t_j / L_jj ; t_i -= L_ij y_j  (i > j)
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
                    dl[s2] = (s2 >= s && i > j && i < Tn) ? D[tri(i) + j] : 0.0;
                }
                const double yj = __shfl_sync(0xffffffffu, t[s], jj) * dinvT[j];
                if (lane == jj) t[s] = yj;
#pragma unroll
                for (int s2 = 0; s2 < 4; ++s2)
                    if (s2 >= s) t[s2] = fma(-dl[s2], yj, t[s2]);
            }
        }
    }
    // backward, row oriented on L' : x_j = t_j / L_jj ; t_k -= L_jk x_j  (k < j)
#pragma unroll
    for (int s = 3; s >= 0; --s) {
        if (32 * s < Tn) {
            const int jend = (Tn - 32 * s) < 32 ? (Tn - 32 * s) : 32;
            for (int jj = jend - 1; jj >= 0; --jj) {
                const int j = 32 * s + jj;
                const double* __restrict__ row = D + tri(j);
                double dl[4];
#pragma unroll
                for (int s2 = 0; s2 < 4; ++s2) {
                    const int k = lane + 32 * s2;
                    dl[s2] = (s2 <= s && k < j) ? row[k] : 0.0;
                }
                const double xj = __shfl_sync(0xffffffffu, t[s], jj) * dinvT[j];
                if (lane == jj) t[s] = xj;
#pragma unroll
                for (int s2 = 0; s2 < 4; ++s2)
                    if (s2 <= s) t[s2] = fma(-dl[s2], xj, t[s2]);
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s)
        if (lane + 32 * s < Tn) yt[lane + 32 * s] = t[s];
}

// ---- dense tail, GRID team (one large instance: the ~2000-bus network) --------------------------------------------
// The top of the elimination tree of a 2000-bus network is a chain of ~600 single-column levels (~740 columns): level-
// scheduled it costs two grid barriers per column and phase.  Here it is one packed dense matrix in global memory
// (2-4 MB: L2-resident), factorised right-looking in panels of 32 columns by the WHOLE grid:
//   A. every CTA factorises the 32 x 32 diagonal block redundantly in its shared memory (no grid barrier; block 0 stores it),
//   B. the rows below the block are forward-substituted against it, one row per thread across the grid,     -- grid.sync
//   C. the trailing matrix gets its rank-32 update in 32 x 32 tiles, one tile per CTA and trip,              -- grid.sync
// i.e. 2 grid barriers per 32 columns instead of 128+.  Tp = Tpad is a multiple of 32 (identity padding).
__device__ inline double dense_factor_grid(GridTeam& T, double* __restrict__ D, double* __restrict__ dinvT, int Tn, int Tp, double* sm) {
    const int tid = threadIdx.x, nth = blockDim.x;
    double* A = sm;                       // [32][33] diagonal block / X_i tile
    double* Bk = sm + GD_NB * GD_LD;      // [32][33] X_k tile
    double* dv = sm + 2 * GD_NB * GD_LD;  // [32] inverse pivots of the block
    double bad = 0.0;
    for (int j0 = 0; j0 < Tp; j0 += GD_NB) {
        // ---- A: diagonal block, redundantly per CTA: a packed copy is factorised by the CTA team's panel-of-4 code
        //      (dense_factor: 2 barriers per 4 columns instead of 3 per column), then laid out square for steps B and C ----
        double* Pk = Bk;  // 528 of its 1056 doubles
        for (int e = tid; e < GD_NB * GD_NB; e += nth) {
            const int r = e >> 5, c = e & 31;
            if (c <= r) Pk[tri(r) + c] = D[tri(j0 + r) + j0 + c];
        }
        __syncthreads();
        bad = fmax(bad, dense_factor(Pk, dv, GD_NB));  // ends with a barrier; dv = inverse pivots
        for (int e = tid; e < GD_NB * GD_NB; e += nth) {
            const int r = e >> 5, c = e & 31;
            A[r * GD_LD + c] = (c < r) ? Pk[tri(r) + c] : (c == r ? 1.0 / dv[r] : 0.0);
        }
        __syncthreads();
        if (blockIdx.x == 0) {
            for (int e = tid; e < GD_NB * GD_NB; e += nth) {
                const int r = e >> 5, c = e & 31;
                if (c <= r) D[tri(j0 + r) + j0 + c] = A[r * GD_LD + c];
            }
            if (tid < GD_NB && j0 + tid < Tn) dinvT[j0 + tid] = dv[tid];
        }
        // ---- B: rows below the block, one per thread across the grid: X_i = A_i L_d^{-T} ----
        {   // a warp per row across the grid, lane c owns x_c: column-oriented substitution with shuffle broadcasts
            const int lane = tid & 31, gw = T.tid() >> 5, nw = T.size() >> 5;
            for (int i = j0 + GD_NB + gw; i < Tp; i += nw) {
                double* __restrict__ row = D + tri(i) + j0;
                double t = row[lane];
                const double di = dv[lane];
                for (int k = 0; k < GD_NB; ++k) {
                    const double xk = __shfl_sync(0xffffffffu, t * di, k);
                    if (lane == k) t = xk;
                    else if (lane > k) t = fma(-xk, A[lane * GD_LD + k], t);
                }
                row[lane] = t;
            }
        }
        T.sync();
        // ---- C: rank-32 update of the trailing matrix in 32 x 32 tiles ----
        const int nb = (Tp - j0 - GD_NB) / GD_NB;
        const int ntiles = nb * (nb + 1) / 2;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            int bi = 0;
            while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;  // tile (bi, bk), bk <= bi (nb <= ~60: a short scan)
            const int bk = t - bi * (bi + 1) / 2;
            const int i0 = j0 + GD_NB + bi * GD_NB, k0 = j0 + GD_NB + bk * GD_NB;
            __syncthreads();  // previous tile (and step A/B reads of A) done
            for (int e = tid; e < GD_NB * GD_NB; e += nth) {
                const int r = e >> 5, c = e & 31;
                A[r * GD_LD + c] = D[tri(i0 + r) + j0 + c];
                Bk[r * GD_LD + c] = D[tri(k0 + r) + j0 + c];
            }
            __syncthreads();
            for (int e = tid; e < GD_NB * GD_NB; e += nth) {
                const int r = e >> 5, k = e & 31;
                if (i0 + r >= k0 + k) {
                    double sacc = 0.0;
#pragma unroll
                    for (int c = 0; c < GD_NB; ++c) sacc = fma(A[r * GD_LD + c], Bk[k * GD_LD + c], sacc);
                    D[tri(i0 + r) + k0 + k] -= sacc;
                }
            }
        }
        T.sync();
    }
    return bad;  // uniform: every CTA factorised every diagonal block
}

// Tail solves of the grid team, executed by ONE CTA (the others wait at the grid barrier that follows): blocked by 32 and
// LEFT-looking -- the right-hand side of a block is corrected by one mat-vec over everything solved so far (32 rows x j0
// columns, eight lanes per row walking the packed row contiguously, all loads of a lane independent), then one warp solves
// the 32 x 32 diagonal block from shared memory (registers + shuffles).  The right-looking variant it replaces updated every
// remaining row after each block with four dependent 32-load chains per thread: 10 k cycles per block step against ~4 k.
__device__ inline void dense_solve_block0(const double* __restrict__ D, const double* __restrict__ dinvT, double* yt, int Tn, int Tp, double* sm) {
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31;
    double* A = sm;                            // [32][33] diagonal block
    double* red = sm + GD_NB * GD_LD;          // [8][32] partial sums of the backward mat-vec / [32] of the forward one
    for (int j0 = 0; j0 < Tp; j0 += GD_NB) {   // forward: y = L^{-1} y
        __syncthreads();
        {   // rows j0 .. j0+31 against y[0 .. j0): thread = (row r, lane8)
            const int r = tid >> 3, l8 = tid & 7;
            double a0 = 0.0, a1 = 0.0;
            if (r < GD_NB) {
                const double* __restrict__ row = D + tri(j0 + r);
                int k = l8;
#pragma unroll 4
                for (; k + 8 < j0; k += 16) { a0 = fma(row[k], yt[k], a0); a1 = fma(row[k + 8], yt[k + 8], a1); }
                if (k < j0) a0 = fma(row[k], yt[k], a0);
            }
            a0 += a1;
            a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
            a0 += __shfl_xor_sync(0xffffffffu, a0, 2);
            a0 += __shfl_xor_sync(0xffffffffu, a0, 4);
            if (r < GD_NB && l8 == 0) red[r] = a0;
        }
        for (int e = tid; e < GD_NB * GD_NB; e += nth) {
            const int r = e >> 5, c = e & 31;
            A[r * GD_LD + c] = (c <= r) ? D[tri(j0 + r) + j0 + c] : 0.0;
        }
        __syncthreads();
        if (tid < 32) {
            double t = (j0 + lane < Tn) ? yt[j0 + lane] - red[lane] : 0.0;
            const double di = (j0 + lane < Tn) ? dinvT[j0 + lane] : 1.0;
            for (int c = 0; c < GD_NB; ++c) {
                const double yc = __shfl_sync(0xffffffffu, t * di, c);
                if (lane == c) t = yc;
                else if (lane > c) t = fma(-A[lane * GD_LD + c], yc, t);
            }
            if (j0 + lane < Tn) yt[j0 + lane] = t;
        }
    }
    for (int j0 = Tp - GD_NB; j0 >= 0; j0 -= GD_NB) {  // backward: x = L^{-T} y
        __syncthreads();
        {   // columns j0 .. j0+31 against x[j0+32 .. Tn): thread = (row group g, column c), rows i = j0 + 32 + g, + 8, ...
            const int g = tid >> 5, c = tid & 31, ng = nth >> 5;
            double a0 = 0.0, a1 = 0.0;
            int i = j0 + GD_NB + g;
#pragma unroll 4
            for (; i + ng < Tn; i += 2 * ng) {
                a0 = fma(D[tri(i) + j0 + c], yt[i], a0);
                a1 = fma(D[tri(i + ng) + j0 + c], yt[i + ng], a1);
            }
            if (i < Tn) a0 = fma(D[tri(i) + j0 + c], yt[i], a0);
            red[g * 32 + c] = a0 + a1;
        }
        for (int e = tid; e < GD_NB * GD_NB; e += nth) {
            const int r = e >> 5, c = e & 31;
            A[r * GD_LD + c] = (c <= r) ? D[tri(j0 + r) + j0 + c] : 0.0;
        }
        __syncthreads();
        if (tid < 32) {
            double s = 0.0;
            for (int g = 0; g < (nth >> 5); ++g) s += red[g * 32 + lane];
            double t = (j0 + lane < Tn) ? yt[j0 + lane] - s : 0.0;
            const double di = (j0 + lane < Tn) ? dinvT[j0 + lane] : 1.0;
            for (int c = GD_NB - 1; c >= 0; --c) {
                const double xc = __shfl_sync(0xffffffffu, t * di, c);
                if (lane == c) t = xc;
                else if (lane < c) t = fma(-A[c * GD_LD + lane], xc, t);
            }
            if (j0 + lane < Tn) yt[j0 + lane] = t;
        }
    }
    __syncthreads();
}

// team dispatch of the dense-tail code
__device__ __forceinline__ double team_dense_factor(CtaTeam&, const CholDev& C, const CholWork& W) {
    return dense_factor(W.D, W.dinv + C.n0, C.T);
}
__device__ __forceinline__ double team_dense_factor(GridTeam& T, const CholDev& C, const CholWork& W) {
    return dense_factor_grid(T, W.D, W.dinv + C.n0, C.T, C.Tpad, W.gsm);
}
__device__ __forceinline__ void team_dense_solve(CtaTeam&, const CholDev& C, const CholWork& W, double* yw) {
    if (threadIdx.x < 32) dense_solve_warp(W.D, W.dinv + C.n0, yw + C.n0, C.T);
}
__device__ __forceinline__ void team_dense_solve(GridTeam&, const CholDev& C, const CholWork& W, double* yw) {
    if (blockIdx.x == 0) dense_solve_block0(W.D, W.dinv + C.n0, yw + C.n0, C.T, C.Tpad, W.gsm);
}

// In-place numeric factorisation: the barrier phases of symbolic.hpp 4c (per sparse level the
// diagonal entries, then the sub-diagonal entries; then the Schur complement of the dense tail).  A
// thread executes one SLOT per round: one lane's share of a task, the lane count chosen per task on
// the host from its pair count.  Returns false (uniformly) if a pivot was not positive.
// With a fused program (C.fused_fwd) and a right-hand side b, the forward sweep of the solve K x = b rides in the same
// phases (sweep slots: bit 31, pairs (L value, row of yw)); chol_solve is then called with skip_fwd.
template <class Team>
__device__ bool chol_factor(Team& T, const CholDev& C, const CholWork& W, Prof& pf, const double* b = nullptr) {
    double* L = W.L;
    double* yw = W.yw;
    double bad[1] = {0.0};
    if (C.fused_fwd && b)  // first read of yw is in the second phase, behind the barrier of the first
        for (int k = T.tid(); k < C.n; k += T.size()) yw[k] = b[C.perm[k]];
    int4 ph = C.fphase[0];
    for (int p = 0; p < C.nphase; ++p) {
        const int4 nxt = C.fphase[p + 1];  // (padded by one entry) off the critical path of the next phase
        const int s0 = ph.x, ns = ph.y - ph.x, kind = ph.w;
        for (int r0 = 0; r0 < ns; r0 += 2 * T.size()) {  // two slots per thread and trip (gather_dot2)
            const int sA = r0 + T.tid(), sB = sA + T.size();
            const bool onA = sA < ns, onB = sB < ns;
            const int4 slA = C.ftask[s0 + (onA ? sA : 0)], slB = C.ftask[s0 + (onB ? sB : 0)];
            const int eA = slA.x & 0x3ffffff, eB = slB.x & 0x3ffffff;
            const int LnA = 1 << ((slA.x >> 26) & 7), LnB = 1 << ((slB.x >> 26) & 7);
            const bool ldA = onA && ((slA.x >> 29) & 1), ldB = onB && ((slB.x >> 29) & 1);
            const bool swA = slA.x < 0, swB = slB.x < 0;  // forward-sweep slot
            const double kA = !ldA ? 0.0 : (swA ? yw[eA] : (((slA.x >> 30) & 1) ? L[eA] : 0.0));
            const double kB = !ldB ? 0.0 : (swB ? yw[eB] : (((slB.x >> 30) & 1) ? L[eB] : 0.0));
            const double dA = (ldA && kind == 1) ? W.dinv[slA.w] : 0.0, dB = (ldB && kind == 1) ? W.dinv[slB.w] : 0.0;
            double accA, accB;
            gather_dot2(C.fp_ab, onA ? slA.y : 0, onA ? slA.z : 0, LnA, onB ? slB.y : 0, onB ? slB.z : 0, LnB, L, swA ? yw : L,
                        swB ? yw : L, accA, accB);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {  // lane groups of mixed (power of two, aligned) sizes share the butterfly
                const double tA = __shfl_xor_sync(0xffffffffu, accA, o), tB = __shfl_xor_sync(0xffffffffu, accB, o);
                if (o < LnA) accA += tA;
                if (o < LnB) accB += tB;
            }
            if (ldA) {
                double v = kA - accA;
                if (swA) {
                    yw[eA] = kind == 1 ? v * dA : v;  // row of a sparse level / right-hand side of the tail
                } else if (kind == 0) {
                    if (!(v > 0.0)) { bad[0] = 1.0; v = 1.0; }
                    const double inv = rsqrt(v);
                    L[eA] = v * inv;
                    W.dinv[slA.w] = inv;
                } else if (kind == 1) {
                    L[eA] = v * dA;
                } else {
                    W.D[slA.w] = v;
                }
            }
            if (ldB) {
                double v = kB - accB;
                if (swB) {
                    yw[eB] = kind == 1 ? v * dB : v;
                } else if (kind == 0) {
                    if (!(v > 0.0)) { bad[0] = 1.0; v = 1.0; }
                    const double inv = rsqrt(v);
                    L[eB] = v * inv;
                    W.dinv[slB.w] = inv;
                } else if (kind == 1) {
                    L[eB] = v * dB;
                } else {
                    W.D[slB.w] = v;
                }
            }
        }
        T.sync();
        ph = nxt;
    }
    pf.lap(PS_FACTOR_SPARSE);
    if (C.T > 0) {
        bad[0] = fmax(bad[0], team_dense_factor(T, C, W));
        pf.lap(PS_FACTOR_DENSE);
    }
    T.template reduce<1, true>(bad);
    return bad[0] == 0.0;
}

// x = K^{-1} b   (b, x in original order; x may alias b)
// skip_fwd: the forward sweep and the tail right-hand side were done by chol_factor (fused program): yw already holds them
template <class Team>
__device__ void chol_solve(Team& T, const CholDev& C, const CholWork& W, const double* b, double* x, Prof& pf, const bool skip_fwd = false) {
    const double* L = W.L;
    double* yw = W.yw;
    const double* dinv = W.dinv;
    if (!skip_fwd) {
    for (int k = T.tid(); k < C.n; k += T.size()) yw[k] = b[C.perm[k]];
    T.sync();
    }
    for (int l = 0; l < (skip_fwd ? 0 : C.nlev); ++l) {  // forward: rows of L
        const int c0 = C.lev_ptr[l], cnt = C.lev_ptr[l + 1] - c0, lg = level_lg(T, cnt), Ln = 1 << lg;
        const int lane = T.tid() & (Ln - 1), sub = T.tid() >> lg, nsub = T.size() >> lg;
        for (int t0 = 0; t0 < cnt; t0 += nsub) {
            const int t = t0 + sub;
            const bool on = t < cnt;
            const int j = c0 + (on ? t : 0);
            const double yj = yw[j], dj = dinv[j];
            double acc = on ? gather_dot(C.Rci, C.Rp[j] + lane, C.Rp[j + 1], Ln, L, yw) : 0.0;
            acc = subwarp_sum(acc, Ln);
            if (on && lane == 0) yw[j] = (yj - acc) * dj;
        }
        T.sync();
    }
    pf.lap(PS_FWD);
    if (C.T > 0) {
        // tail right-hand side: b_j - sum_{k < n0} L_jk y_k, then the dense solves (one warp)
        const int lg = level_lg(T, C.T), Ln = 1 << lg;
        const int lane = T.tid() & (Ln - 1), sub = T.tid() >> lg, nsub = T.size() >> lg;
        for (int t0 = 0; t0 < (skip_fwd ? 0 : C.T); t0 += nsub) {
            const int t = t0 + sub;
            const bool on = t < C.T;
            const int j = C.n0 + (on ? t : 0);
            double acc = on ? gather_dot(C.Rci, C.Rp[j] + lane, C.Rmid[j], Ln, L, yw) : 0.0;
            acc = subwarp_sum(acc, Ln);
            if (on && lane == 0) yw[j] -= acc;
        }
        T.sync();
        team_dense_solve(T, C, W, yw);
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
            const double yj = yw[j], dj = dinv[j];
            double acc = on ? column_dot(L, C.Li, C.Lp[j] + 1 + lane, C.Lp[j + 1], Ln, yw) : 0.0;
            acc = subwarp_sum(acc, Ln);
            if (on && lane == 0) yw[j] = (yj - acc) * dj;
        }
        T.sync();
    }
    for (int k = T.tid(); k < C.n; k += T.size()) x[C.perm[k]] = yw[k];
    T.sync();
    pf.lap(PS_BWD);
}

// ---- resident CTA team: the same factorisation and solves driven by the ring program ---------------------------------------
// Assembly (wJ pre-pass and tail clear as in chol_assemble, then the assembly chunks), the sparse levels WITH the forward
// sweep of right-hand side b riding in their chunks, and the Schur complement with the tail right-hand side are ONE segment,
// so the stream never stops between them; the backward sweep is requested while the dense tail is being factorised.
// The tail block of L is not used in this mode: its assembled K goes straight into D.
__device__ __forceinline__ bool chol_assemble_factor_fwd_ring(CtaTeam& T, Ring& R, const CholDev& C, const CholWork& W,
                                                              const double* __restrict__ Pv, const double* __restrict__ dg,
                                                              const double shift, const double* __restrict__ w,
                                                              const double* __restrict__ Jv, const double* __restrict__ b, Prof& pf) {
    ring_prime(R, C, 0);  // in flight while wJ is formed
    double* wJ = W.wJ;
    for (int a0 = T.tid(); a0 < C.nslotJ; a0 += 4 * T.size()) {
        int r[4];
        double jv[4], wr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int a = a0 + u * T.size();
            const bool ok = a < C.nslotJ;
            r[u] = ok ? C.jrow[a] : -1;
            jv[u] = ok ? Jv[a] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) wr[u] = r[u] >= 0 ? w[r[u]] : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int a = a0 + u * T.size();
            if (a < C.nslotJ) wJ[a] = wr[u] * jv[u];
        }
    }
    for (int k = T.tid(); k < C.n; k += T.size()) W.yw[k] = b[C.perm[k]];  // a failed attempt has overwritten it
    if (C.T > 0) {
        const int Tp = C.Tpad;
        for (int i = T.tid(); i < Tp * (Tp + 1) / 2; i += T.size()) W.D[i] = 0.0;
        T.sync();
        for (int i = C.T + T.tid(); i < Tp; i += T.size()) W.D[i * (i + 1) / 2 + i] = 1.0;
    }
    T.sync();
    pf.lap(PS_ASSEMBLE);
    RingArgs a{Pv, dg, shift, Jv};
    double bad[1];
    bad[0] = ring_run_segment(R, C, W, 0, 2, a, pf);
    pf.lap(PS_FACTOR_SPARSE);
    if (C.T > 0) {
        bad[0] = fmax(bad[0], dense_factor(W.D, W.dinv + C.n0, C.T));
        pf.lap(PS_FACTOR_DENSE);
    }
    T.template reduce<1, true>(bad);
    return bad[0] == 0.0;
}

// the rest of the solve after chol_assemble_factor_fwd_ring: dense tail, backward sweep, x in original order
__device__ __forceinline__ void chol_backsolve_ring(CtaTeam& T, Ring& R, const CholDev& C, const CholWork& W, double* x, Prof& pf) {
    double* yw = W.yw;
    if (C.T > 0) {
        if (threadIdx.x < 32) dense_solve_warp(W.D, W.dinv + C.n0, yw + C.n0, C.T);
        T.sync();
        pf.lap(PS_TAIL);
    }
    RingArgs a{nullptr, nullptr, 0.0, nullptr};
    ring_run_segment(R, C, W, 2, 0, a, pf);
    for (int k = T.tid(); k < C.n; k += T.size()) x[C.perm[k]] = yw[k];
    T.sync();
    pf.lap(PS_BWD);
}
