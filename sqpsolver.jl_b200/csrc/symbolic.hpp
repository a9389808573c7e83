// symbolic.hpp -- host-side symbolic analysis of the condensed Newton/KKT matrix
//
//      K = P + diag(d) + J' diag(w) J            (n x n, symmetric positive definite)
//
// shared by every instance of a batch (one sparsity pattern) and by every iteration of a
// solve (only d, w and the values of P, J change).  Done once at setup, in plain C++:
//   1. pattern of K = pattern(P) U pattern(J'J) U I
//   2. fill-reducing ordering: minimum degree on the elimination graph (exact external degree)
//   3. elimination tree and its levels; the columns are then RENUMBERED LEVEL-MAJOR (any
//      topological order of the elimination tree gives the same fill), so that
//        * the columns of one level are a contiguous range -> no indirection through a level list,
//        * the entries of L (CSC, diagonal first) are stored in execution order,
//        * the top of the tree -- a chain of single-column levels, where the factor is nearly
//          dense and a level-scheduled sparse code would pay one barrier and one dependent gather
//          chain per column -- is the LAST T columns: the "dense tail".  Its Schur complement is
//          formed in one parallel phase and factorised as a dense packed matrix in shared memory.
//   4. the index programs the device kernels execute:
//        assembly : K_e  = P[h] + d[diag] + sum_t w[row_t] * Jv[a_t] * Jv[b_t]
//        factor   : L_e  = (K_e - sum_t L[p_t] * L[q_t]) / L_jj    over columns k < min(j, n0)
//        solve    : level-scheduled forward (rows of L) and backward (columns of L) sweeps
// No numerical work happens here.
#pragma once
#include <algorithm>
#include <cstdint>
#include <utility>
#include <vector>

struct Symbolic {
    int n = 0, nnzL = 0;
    int nlev = 0;                        // number of SPARSE levels (columns 0 .. n0-1)
    int n0 = 0, T = 0;                   // dense tail = columns n0 .. n-1 (T = n - n0; 0 = none)
    int nlev_total = 0;                  // levels of the whole elimination tree (statistics)
    std::vector<int> perm, iperm;        // perm[k] = original index of pivot k
    std::vector<int> Lp, Li;             // CSC of L (permuted indices), diagonal first in each column
    std::vector<int> Rp, Rmid;           // CSR of strictly-lower L: row ptr; Rmid[j] = end of the part with column < n0
    std::vector<int> Rci;                // interleaved (index into the CSC value array, column) per CSR entry
    std::vector<int> lev_ptr;            // [nlev+1] column range of each sparse level
    std::vector<int> fp_ptr;             // per entry id: range in fp_ab
    std::vector<int> fp_ab;              // interleaved pairs of L value indices to multiply-subtract
    // factorisation as a list of barrier phases over flat SLOT lists (4 ints each).  A task (one entry of L) gets
    // 2^lg lanes, lg chosen from its pair count (<= 8 pairs per lane, widened when the phase is narrow); every lane
    // of a task has its own slot, so a thread needs ONE coalesced 16-byte load to know its work:
    //   slot  = (entry id | lg << 26 | leader << 29 | has_K << 30, first pair of this lane, end pair, aux)
    //           the lane walks its pairs with stride 2^lg; aux: column (diag/offdiag) or packed tail position (Schur)
    //   phase = (first slot, end slot, max pairs of a task, kind)     kind: 0 diagonal entries of a level, 1 sub-diagonal
    //           entries of a level, 2 Schur complement of the dense tail.   Tasks of a phase are sorted by pair count
    //           (descending): lane groups are then aligned to their own (power of two) size inside a warp.
    std::vector<int> ftask, fphase;
    int ftasks = 0;                      // number of tasks (entries) behind the slots
    // fused forward sweep (fuse_fwd): the right-hand side of the Newton solve is known before the factorisation, and row j of
    // the forward sweep needs what the sub-diagonal entries of column j need (the pivot of j, finished lower levels), so its
    // tasks ride in the same phase -- one barrier phase per level saved; the tail right-hand side rides with the Schur
    // complement.  A sweep slot has bit 31 set, target = row j, aux = j, and its pairs (value index, row) are the CSR entries
    // of row j, appended to fp_ab at pair offset sw_off.
    bool fused_fwd = false;
    int sw_off = 0;
    // assembly: K_e = P[h] + sum_t wJ[a_t] * Jv[b_t]  (+ d[perm[j]] on the diagonal), wJ = w[row] .* Jv
    // assembly slots, same lane-group scheme as the factor slots (a diagonal entry of a bus variable has 40+ terms, most
    // sourced entries have one or two):  slot = (entry id | lg << 26 | leader << 29, first term of this lane, end term,
    // P value index | -1);  aslot_d[slot] = original column whose d[] goes to this (diagonal) entry, or -1
    std::vector<int> aslot, aslot_d;
    int atasks = 0;
    std::vector<int> as_ab;              // interleaved (Jv index a, Jv index b) per term
    std::vector<int> jrow;               // row of every J value slot (for wJ); slots beyond a row's end: -1
    int as_terms = 0;
    int64_t flops = 0;                   // of the sparse part (2 * pairs)
    bool ok = false;                     // false: a row of J is too long for the clique expansion
    // kept for build_slot_programs(): per entry the P value index (-1: none), the range of its assembly terms, has_K
    std::vector<int> as_h, as_ptr;
    std::vector<char> hasK;
};

// The slot lists of the assembly and of the factorisation phases for ONE team shape: a task gets 2^lg lanes with
// lg <= lgmax (the lanes of a task must sit in one warp: 32 lanes for a CTA team, 32 / G task lanes when G instances
// are interleaved in a warp), and narrow phases are widened while one round of `team_lanes` lanes can hold them.
struct SlotProg {
    std::vector<int> ftask, fphase, aslot, aslot_d;
    int ftasks = 0, atasks = 0;
};
inline void build_slot_programs(const Symbolic& S, int lgmax, int team_lanes, SlotProg& out, bool fuse_fwd = false);

// J: m x ncols CSR (rb/re per row, so a prefix of each row can be used), P: symmetric-full CSR or null.
// tail_max: largest dense tail (columns) the caller can hold; 0 disables the dense tail.
inline Symbolic symbolic_analyze(int n, int m, const int* Jrb, const int* Jre, const int* Jcol, const int* Prp,
                                 const int* Pcol, int max_row_len = 512, int tail_max = 0, bool fuse_fwd = false) {
    Symbolic S;
    S.n = n;
    // ---- 1. adjacency of K ---------------------------------------------------------------
    std::vector<std::vector<int>> adj(n);
    for (int i = 0; i < m; ++i) {
        int len = Jre[i] - Jrb[i];
        if (len > max_row_len) return S;
        for (int a = Jrb[i]; a < Jre[i]; ++a)
            for (int b = Jrb[i]; b < Jre[i]; ++b)
                if (Jcol[a] != Jcol[b]) adj[Jcol[a]].push_back(Jcol[b]);
    }
    if (Prp)
        for (int j = 0; j < n; ++j)
            for (int k = Prp[j]; k < Prp[j + 1]; ++k)
                if (Pcol[k] != j) { adj[j].push_back(Pcol[k]); adj[Pcol[k]].push_back(j); }
    for (auto& v : adj) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }

    // ---- 2. minimum degree elimination; column structures fall out ---------------------------
    std::vector<char> done(n, 0);
    std::vector<int> order;
    order.reserve(n);
    std::vector<std::vector<int>> colstruct(n);  // by pivot position: neighbours (original ids) at elimination time
    std::vector<int> mark(n, -1), tmp;
    for (int k = 0; k < n; ++k) {
        int best = -1, bestdeg = 1 << 30;
        for (int v = 0; v < n; ++v)
            if (!done[v] && (int)adj[v].size() < bestdeg) { bestdeg = (int)adj[v].size(); best = v; }
        int v = best;
        done[v] = 1;
        order.push_back(v);
        colstruct[k] = adj[v];
        // neighbours become a clique, v disappears
        for (int u : adj[v]) {
            tmp.clear();
            for (int w : adj[u]) if (w != v) { mark[w] = u; tmp.push_back(w); }
            for (int w : adj[v]) if (w != u && mark[w] != u) tmp.push_back(w);
            std::sort(tmp.begin(), tmp.end());
            adj[u].swap(tmp);
        }
        adj[v].clear();
        adj[v].shrink_to_fit();
    }
    // ---- 3. elimination tree, levels, level-major renumbering, dense tail ---------------------
    {
        std::vector<int> ip0(n), parent(n, -1), level(n, 0);
        for (int k = 0; k < n; ++k) ip0[order[k]] = k;
        for (int k = 0; k < n; ++k) {
            int p = n;
            for (int u : colstruct[k]) p = std::min(p, ip0[u]);
            parent[k] = p < n ? p : -1;
        }
        for (int k = 0; k < n; ++k)  // parent[k] > k: one ascending pass settles every level
            if (parent[k] >= 0) level[parent[k]] = std::max(level[parent[k]], level[k] + 1);
        int nl = 0;
        for (int k = 0; k < n; ++k) nl = std::max(nl, level[k] + 1);
        S.nlev_total = nl;
        std::vector<int> pos(n);
        for (int k = 0; k < n; ++k) pos[k] = k;
        std::stable_sort(pos.begin(), pos.end(), [&](int a, int b) { return level[a] < level[b]; });
        std::vector<int> cnt(nl + 1, 0);
        for (int k = 0; k < n; ++k) cnt[level[k] + 1]++;
        for (int l = 0; l < nl; ++l) cnt[l + 1] += cnt[l];
        // dense tail: the top levels, as many as fit tail_max columns
        int lcut = nl;
        while (lcut > 0 && n - cnt[lcut - 1] <= tail_max) --lcut;
        int T = n - cnt[lcut];
        if (T < 16 || lcut == 0) { lcut = nl; T = 0; }  // not worth it / everything dense (tiny problems): plain sparse code
        S.T = T;
        S.n0 = n - T;
        S.nlev = lcut;
        S.lev_ptr.assign(cnt.begin(), cnt.begin() + lcut + 1);
        std::vector<int> order2(n);
        std::vector<std::vector<int>> cs2(n);
        for (int k = 0; k < n; ++k) { order2[k] = order[pos[k]]; cs2[k].swap(colstruct[pos[k]]); }
        order.swap(order2);
        colstruct.swap(cs2);
    }
    S.perm = order;
    S.iperm.assign(n, 0);
    for (int k = 0; k < n; ++k) S.iperm[order[k]] = k;
    // CSC of L in permuted indices, diagonal first then ascending rows
    S.Lp.assign(n + 1, 0);
    for (int k = 0; k < n; ++k) S.Lp[k + 1] = S.Lp[k] + 1 + (int)colstruct[k].size();
    S.nnzL = S.Lp[n];
    S.Li.resize(S.nnzL);
    for (int k = 0; k < n; ++k) {
        int p = S.Lp[k];
        S.Li[p++] = k;
        std::vector<int> rows;
        for (int u : colstruct[k]) rows.push_back(S.iperm[u]);
        std::sort(rows.begin(), rows.end());
        for (int r : rows) S.Li[p++] = r;
    }
    // CSR of strictly-lower L (row i: columns k < i, ascending)
    const int n0 = S.n0;
    std::vector<int> Rc, Ri;
    S.Rp.assign(n + 1, 0);
    for (int k = 0; k < n; ++k)
        for (int p = S.Lp[k] + 1; p < S.Lp[k + 1]; ++p) S.Rp[S.Li[p] + 1]++;
    for (int i = 0; i < n; ++i) S.Rp[i + 1] += S.Rp[i];
    Rc.resize(S.Rp[n]);
    Ri.resize(S.Rp[n]);
    {
        std::vector<int> cur(S.Rp.begin(), S.Rp.end() - 1);
        for (int k = 0; k < n; ++k)
            for (int p = S.Lp[k] + 1; p < S.Lp[k + 1]; ++p) {
                int i = S.Li[p];
                Rc[cur[i]] = k;
                Ri[cur[i]++] = p;
            }
    }
    S.Rmid.resize(n);
    S.Rci.resize(2 * (size_t)S.Rp[n]);
    for (int i = 0; i < n; ++i) {
        int q = S.Rp[i];
        while (q < S.Rp[i + 1] && Rc[q] < n0) ++q;
        S.Rmid[i] = q;
    }
    for (int q = 0; q < S.Rp[n]; ++q) { S.Rci[2 * q] = Ri[q]; S.Rci[2 * q + 1] = Rc[q]; }
    // ---- 4a. factorisation program ------------------------------------------------------------
    // entry (i,j), j<=i:  sum over k < min(j, n0) with L_ik != 0 and L_jk != 0 -> intersect rows i and j of the CSR
    S.fp_ptr.assign(S.nnzL + 1, 0);
    for (int pass = 0; pass < 2; ++pass) {
        int64_t total = 0;
        for (int j = 0; j < n; ++j) {
            const int lim = std::min(j, n0);
            for (int p = S.Lp[j]; p < S.Lp[j + 1]; ++p) {
                int i = S.Li[p];
                int a = S.Rp[i], ae = S.Rp[i + 1], b = S.Rp[j], be = S.Rp[j + 1];
                int cnt = 0;
                while (a < ae && b < be) {
                    int ca = Rc[a], cb = Rc[b];
                    if (ca >= lim || cb >= lim) break;
                    if (ca == cb) {
                        if (pass) { S.fp_ab[2 * (size_t)(S.fp_ptr[p] + cnt)] = Ri[a]; S.fp_ab[2 * (size_t)(S.fp_ptr[p] + cnt) + 1] = Ri[b]; }
                        ++cnt; ++a; ++b;
                    } else if (ca < cb) ++a; else ++b;
                }
                if (!pass) S.fp_ptr[p + 1] = cnt;
                total += cnt;
            }
        }
        if (!pass) {
            for (int p = 0; p < S.nnzL; ++p) S.fp_ptr[p + 1] += S.fp_ptr[p];
            S.fp_ab.resize(2 * (size_t)S.fp_ptr[S.nnzL]);
            S.flops = 2 * total;
        }
    }
    // ---- 4b. assembly program --------------------------------------------------------------------
    // map (i,j) permuted, i>=j -> entry id
    std::vector<int> as_h(S.nnzL, -1);
    auto find_entry = [&](int a, int b) {  // original indices
        int i = S.iperm[a], j = S.iperm[b];
        if (i < j) std::swap(i, j);
        const int* lo = &S.Li[S.Lp[j]] + 1;
        const int* hi = &S.Li[0] + S.Lp[j + 1];
        if (i == j) return S.Lp[j];
        const int* it = std::lower_bound(lo, hi, i);
        return (int)(it - &S.Li[0]);
    };
    if (Prp)
        for (int a = 0; a < n; ++a)
            for (int k = Prp[a]; k < Prp[a + 1]; ++k) {
                int b = Pcol[k];
                if (S.iperm[a] >= S.iperm[b]) as_h[find_entry(a, b)] = k;  // lower triangle in permuted order
            }
    std::vector<int> cnt(S.nnzL + 1, 0), as_ptr;
    for (int pass = 0; pass < 2; ++pass) {
        for (int r = 0; r < m; ++r)
            for (int a = Jrb[r]; a < Jre[r]; ++a)
                for (int b = Jrb[r]; b < Jre[r]; ++b) {
                    int ca = Jcol[a], cb = Jcol[b];
                    if (S.iperm[ca] < S.iperm[cb]) continue;  // keep i >= j once
                    if (ca == cb && a != b) continue;         // (no duplicate columns in a CSR row)
                    int e = find_entry(ca, cb);
                    if (!pass) cnt[e + 1]++;
                    else {
                        size_t pos = (size_t)as_ptr[e] + cnt[e]++;
                        S.as_ab[2 * pos] = a; S.as_ab[2 * pos + 1] = b;
                    }
                }
        if (!pass) {
            as_ptr.assign(S.nnzL + 1, 0);
            for (int e = 0; e < S.nnzL; ++e) as_ptr[e + 1] = as_ptr[e] + cnt[e + 1];
            S.as_ab.resize(2 * (size_t)as_ptr[S.nnzL]);
            S.as_terms = as_ptr[S.nnzL];
            std::fill(cnt.begin(), cnt.end(), 0);
        }
    }
    S.as_h = as_h;
    S.as_ptr = as_ptr;
    S.hasK.assign(S.nnzL, 0);
    for (int j = 0; j < n; ++j)
        for (int e = S.Lp[j]; e < S.Lp[j + 1]; ++e) {
            const bool diag = e == S.Lp[j];
            if (!diag && as_h[e] < 0 && as_ptr[e] == as_ptr[e + 1]) continue;  // pure fill: K_e = 0, nothing to assemble
            S.hasK[e] = 1;
        }
    {
        int nslots = 0;
        for (int r = 0; r < m; ++r) nslots = std::max(nslots, Jre[r]);
        for (int r = 0; r < m; ++r) nslots = std::max(nslots, Jrb[r]);
        S.jrow.assign(nslots, -1);
        for (int r = 0; r < m; ++r)
            for (int a = Jrb[r]; a < Jre[r]; ++a) S.jrow[a] = r;
    }
    // ---- 4c. slot lists of the assembly and the factorisation phases (CTA team: 32 lanes per task at most) ----------
    {
        if (fuse_fwd) {  // the sweep tasks take their pairs from the CSR of L, appended behind the factor pairs
            S.fused_fwd = true;
            S.sw_off = (int)(S.fp_ab.size() / 2);
            S.fp_ab.insert(S.fp_ab.end(), S.Rci.begin(), S.Rci.end());
        }
        SlotProg sp;
        build_slot_programs(S, 5, 512, sp, fuse_fwd);
        S.ftask.swap(sp.ftask); S.fphase.swap(sp.fphase); S.aslot.swap(sp.aslot); S.aslot_d.swap(sp.aslot_d);
        S.ftasks = sp.ftasks; S.atasks = sp.atasks;
    }
    S.ok = true;
    return S;
}

// ---- ring programs: the same index work as CHUNK IMAGES for the shared-memory ring of the resident CTA team (chol.cuh) ----
// When one CTA owns an SM and keeps the factor, the solve scratch and the gathered vectors in shared memory, the index
// program of every barrier phase (slot -> pair list) is the only thing left to fetch.  It is static, so it is laid out here
// as a sequence of self-contained chunk images that one thread streams into a ring of shared-memory stages with bulk
// asynchronous copies (cp.async.bulk + mbarrier) a few chunks ahead of the consumers.  One chunk = up to `ns_max` slots of ONE
// barrier phase:
//   header  8 words : [0] npad (slots rounded up to 32)  [1] kmax (pairs per slot, even)  [2] 1 = assembly chunk (kinds 5 / 6:
//                     operands in global memory, pair counts honoured), 0 = factor / sweep chunk (operands in shared memory, short
//                     pair lists padded with the zero entry L[nL])  [3] last chunk of its segment
//                     [4] word offset of the chunk `stages` positions later in the segment, or -1  [5] its bytes  [6] slots in use
//   slots   npad x 2: w0 = target (16 bits) | lg << 16 | leader << 19 | has_K << 20 | pairs of this lane << 21 | kind << 28 ;  w1 = aux
//   pairs   kmax x npad words, k-major (thread t reads word k * npad + t: conflict-free): low half = index a, high half = index b
// kinds: 0 diagonal entry of a level (aux = column)   1 sub-diagonal entry (aux = column)   2 Schur complement of the dense
//   tail (aux = packed position)   3 row of a triangular sweep: yw[t] = (yw[t] - sum L[a] yw[b]) dinv[t]
//   4 tail right-hand side: yw[t] -= sum L[a] yw[b]   5 assembly into L[t]   6 assembly into the packed tail D[t]
//   (5 / 6: a = wJ index, b = Jv index, aux = (P value index + 1) | (diagonal column + 1) << 16, 0 = none)
// Segments: 0 = assembly + factorisation WITH the forward sweep fused in (the right-hand side is known before the
//   factorisation in the interior-point iteration, and row j of the forward sweep needs exactly what the sub-diagonal entries
//   of column j need -- the pivot of j and finished lower levels -- so its tasks ride in the same chunks; the tail right-hand
//   side rides with the Schur complement), 1 = forward sweep + tail right-hand side on their own (further solves with the
//   same factor), 2 = backward sweep.
// All indices fit 16 bits or the program is not built (ok = false: the caller keeps the slot lists).  The tail block of L is
// never touched in this mode (its assembled K goes straight into D), so only the first nL = Lp[n0] values (+ the zero entry)
// must be resident.
struct RingProg {
    std::vector<int> words;
    std::vector<int> chunk_off, chunk_len;  // per chunk: word offset / words of its image
    int seg_first[3] = {0, 0, 0}, seg_count[3] = {0, 0, 0};
    int stage_words = 0;                    // largest image
    int nL = 0;
    int nchunks = 0;
    bool ok = false;
};

inline void build_ring_program(const Symbolic& S, bool has_P, int ns_max, int kcap, int stage_bytes, int stages, RingProg& out) {
    out = RingProg();
    const int n = S.n, n0 = S.n0;
    out.nL = S.Lp[n0];
    if (out.nL + 1 >= 65535 || n >= 65535 || (int)S.jrow.size() >= 65535 || S.T * (S.T + 1) / 2 + S.T >= 65535) return;
    struct RT { int tgt, aux, hasK, kind; std::vector<std::pair<int, int>> pr; };
    bool fail = false;
    const int nL = out.nL;
    auto emit = [&](std::vector<RT>& v) {
        if (v.empty()) return;
        const bool is_asm = v[0].kind >= 5;
        std::stable_sort(v.begin(), v.end(), [](const RT& a, const RT& b) { return a.pr.size() > b.pr.size(); });
        std::vector<int> lg(v.size(), 0);
        long total = 0;
        for (size_t t = 0; t < v.size(); ++t) {
            const int np = (int)v[t].pr.size();
            while (lg[t] < 5 && np > (kcap << lg[t])) ++lg[t];
            total += 1 << lg[t];
        }
        while (total * 2 <= ns_max) {  // narrow phase: more lanes per task while one round of the team holds them
            bool any = false;
            total = 0;
            for (size_t t = 0; t < v.size(); ++t) {
                if (lg[t] < 5 && (int)v[t].pr.size() > (1 << lg[t])) { ++lg[t]; any = true; }
                total += 1 << lg[t];
            }
            if (!any) break;
        }
        size_t t = 0;
        while (t < v.size()) {
            int ns = 0, kmax = 0;
            size_t u = t;
            while (u < v.size()) {
                const int L = 1 << lg[u];
                int kp = ((int)v[u].pr.size() + L - 1) / L;
                kp = (kp + 1) & ~1;
                const int ns2 = ns + L, km2 = kp > kmax ? kp : kmax, npad2 = (ns2 + 31) & ~31;
                const long bytes = 32 + (long)npad2 * 8 + (long)km2 * npad2 * 4;
                if (ns2 > ns_max || bytes > stage_bytes || km2 > 126) break;
                ns = ns2; kmax = km2; ++u;
            }
            if (u == t) { fail = true; return; }  // one task alone does not fit a stage
            const int npad = (ns + 31) & ~31;
            const int off = (int)out.words.size();
            const int len = 8 + 2 * npad + kmax * npad;
            out.words.resize(off + len, 0);
            int* W = &out.words[off];
            W[0] = npad; W[1] = kmax; W[2] = is_asm ? 1 : 0; W[3] = 0; W[4] = -1; W[5] = 0; W[6] = ns; W[7] = 0;
            int* sl = W + 8;
            int* pw = W + 8 + 2 * npad;
            if (!is_asm)  // short pair lists and idle slots multiply the zero entry with itself
                for (int q = 0; q < kmax * npad; ++q) pw[q] = (int)((unsigned)nL | ((unsigned)nL << 16));
            int s = 0;
            for (size_t k = t; k < u; ++k) {
                const int L = 1 << lg[k], np = (int)v[k].pr.size();
                const unsigned padb = (v[k].kind == 3 || v[k].kind == 4) ? 0u : (unsigned)nL;  // b indexes yw in the sweeps
                for (int lane = 0; lane < L; ++lane, ++s) {
                    int ks = 0;
                    for (int q = lane; q < np; q += L, ++ks) {
                        const unsigned a = (unsigned)v[k].pr[q].first, b = (unsigned)v[k].pr[q].second;
                        pw[ks * npad + s] = (int)(a | (b << 16));
                    }
                    if (!is_asm)
                        for (int q = ks; q < kmax; ++q) pw[q * npad + s] = (int)((unsigned)nL | (padb << 16));
                    sl[2 * s] = v[k].tgt | (lg[k] << 16) | (lane == 0 ? (1 << 19) : 0) | (v[k].hasK ? (1 << 20) : 0) | (ks << 21) | (v[k].kind << 28);
                    sl[2 * s + 1] = v[k].aux;
                }
            }
            out.chunk_off.push_back(off);
            out.chunk_len.push_back(len);
            if (len > out.stage_words) out.stage_words = len;
            t = u;
        }
        v.clear();
    };
    auto close_segment = [&](int seg, int first) {
        out.seg_first[seg] = first;
        out.seg_count[seg] = (int)out.chunk_off.size() - first;
        for (int c = first; c < (int)out.chunk_off.size(); ++c) {
            int* W = &out.words[out.chunk_off[c]];
            W[3] = (c + 1 == (int)out.chunk_off.size()) ? 1 : 0;
            if (c + stages < (int)out.chunk_off.size()) { W[4] = out.chunk_off[c + stages]; W[5] = out.chunk_len[c + stages] * 4; }
        }
    };
    auto row_task = [&](int j, int qe, int kind) {  // row j of the strictly-lower CSR up to qe
        RT t; t.tgt = j; t.aux = 0; t.hasK = 0; t.kind = kind;
        for (int q = S.Rp[j]; q < qe; ++q) t.pr.emplace_back(S.Rci[2 * (size_t)q], S.Rci[2 * (size_t)q + 1]);
        return t;
    };
    std::vector<RT> v;
    // ---- segment 0: assembly (entries of the sparse columns into L, entries of the tail block into D), factorisation + forward sweep ----
    {
        std::vector<RT> vd;
        for (int j = 0; j < n; ++j)
            for (int e = S.Lp[j]; e < S.Lp[j + 1]; ++e) {
                if (!S.hasK[e]) continue;
                RT t;
                t.hasK = 0;
                const int h = has_P ? S.as_h[e] : -1, d = (e == S.Lp[j]) ? S.perm[j] : -1;
                if (h + 1 >= 65535) { return; }
                t.aux = (h + 1) | ((d + 1) << 16);
                for (int q = S.as_ptr[e]; q < S.as_ptr[e + 1]; ++q) t.pr.emplace_back(S.as_ab[2 * (size_t)q], S.as_ab[2 * (size_t)q + 1]);
                if (j < n0) { t.tgt = e; t.kind = 5; v.push_back(std::move(t)); }
                else { const int r = S.Li[e] - n0, c = j - n0; t.tgt = r * (r + 1) / 2 + c; t.kind = 6; vd.push_back(std::move(t)); }
            }
        emit(v);
        emit(vd);
    }
    auto fpairs = [&](int e, RT& t) {
        for (int q = S.fp_ptr[e]; q < S.fp_ptr[e + 1]; ++q) t.pr.emplace_back(S.fp_ab[2 * (size_t)q], S.fp_ab[2 * (size_t)q + 1]);
    };
    for (int l = 0; l < S.nlev && !fail; ++l) {
        for (int j = S.lev_ptr[l]; j < S.lev_ptr[l + 1]; ++j) { RT t; t.tgt = S.Lp[j]; t.aux = j; t.hasK = S.hasK[S.Lp[j]]; t.kind = 0; fpairs(S.Lp[j], t); v.push_back(std::move(t)); }
        emit(v);
        for (int j = S.lev_ptr[l]; j < S.lev_ptr[l + 1]; ++j) {
            for (int e = S.Lp[j] + 1; e < S.Lp[j + 1]; ++e) { RT t; t.tgt = e; t.aux = j; t.hasK = S.hasK[e]; t.kind = 1; fpairs(e, t); v.push_back(std::move(t)); }
            v.push_back(row_task(j, S.Rp[j + 1], 3));
        }
        emit(v);
    }
    if (S.T > 0 && !fail) {
        for (int j = n0; j < n; ++j) {
            for (int e = S.Lp[j]; e < S.Lp[j + 1]; ++e) {
                RT t; const int r = S.Li[e] - n0, c = j - n0;
                t.tgt = 0; t.aux = r * (r + 1) / 2 + c; t.hasK = 0; t.kind = 2; fpairs(e, t);
                if (!t.pr.empty()) v.push_back(std::move(t));  // D already holds K (or 0): nothing to subtract, nothing to do
            }
            RT t = row_task(j, S.Rmid[j], 4);
            if (!t.pr.empty()) v.push_back(std::move(t));
        }
        emit(v);
    }
    if (fail) return;
    close_segment(0, 0);
    // ---- segment 1: forward sweep (rows of L, level by level), then the right-hand side of the tail ----
    int first = (int)out.chunk_off.size();
    for (int l = 0; l < S.nlev && !fail; ++l) {
        for (int j = S.lev_ptr[l]; j < S.lev_ptr[l + 1]; ++j) v.push_back(row_task(j, S.Rp[j + 1], 3));
        emit(v);
    }
    if (S.T > 0 && !fail) {
        for (int j = n0; j < n; ++j) {
            RT t = row_task(j, S.Rmid[j], 4);
            if (!t.pr.empty()) v.push_back(std::move(t));
        }
        emit(v);
    }
    if (fail) return;
    close_segment(1, first);
    // ---- segment 2: backward sweep (columns of L, levels descending) ----
    first = (int)out.chunk_off.size();
    for (int l = S.nlev - 1; l >= 0 && !fail; --l) {
        for (int j = S.lev_ptr[l]; j < S.lev_ptr[l + 1]; ++j) {
            RT t; t.tgt = j; t.aux = 0; t.hasK = 0; t.kind = 3;
            for (int p2 = S.Lp[j] + 1; p2 < S.Lp[j + 1]; ++p2) t.pr.emplace_back(p2, S.Li[p2]);
            v.push_back(std::move(t));
        }
        emit(v);
    }
    if (fail) return;
    close_segment(2, first);
    out.nchunks = (int)out.chunk_off.size();
    out.ok = out.seg_count[0] > 0 && out.seg_count[1] > 0 && out.seg_count[2] > 0;
}

inline void build_slot_programs(const Symbolic& S, int lgmax, int team_lanes, SlotProg& out, bool fuse_fwd) {
    const int n = S.n, n0 = S.n0;
    out = SlotProg();
    // ---- assembly slots (4b) -------------------------------------------------------------------------------------
    {
        struct At { int e, q0, q1, h, d; };
        std::vector<At> v;
        for (int j = 0; j < n; ++j)
            for (int e = S.Lp[j]; e < S.Lp[j + 1]; ++e) {
                if (!S.hasK[e]) continue;
                v.push_back(At{e, S.as_ptr[e], S.as_ptr[e + 1], S.as_h[e], e == S.Lp[j] ? S.perm[j] : -1});
            }
        std::stable_sort(v.begin(), v.end(), [](const At& a, const At& b) { return a.q1 - a.q0 > b.q1 - b.q0; });
        out.atasks = (int)v.size();
        for (const At& t : v) {
            int lg = 0;
            while (lg < lgmax && (t.q1 - t.q0) > (4 << lg)) ++lg;
            for (int lane = 0; lane < (1 << lg); ++lane) {
                out.aslot.push_back(t.e | (lg << 26) | (lane == 0 ? (1 << 29) : 0));
                out.aslot.push_back(t.q0 + lane); out.aslot.push_back(t.q1); out.aslot.push_back(t.h);
                out.aslot_d.push_back(t.d);
            }
        }
    }
    // ---- factorisation phases (4c) ---------------------------------------------------------------------------------
    struct Tk { int tgt, q0, q1, aux; bool sw = false; };
    auto emit = [&](std::vector<Tk>& v, int kind) {
        std::stable_sort(v.begin(), v.end(), [](const Tk& a, const Tk& b) { return a.q1 - a.q0 > b.q1 - b.q0; });
        std::vector<int> lg(v.size(), 0);
        long total = 0;
        for (size_t t = 0; t < v.size(); ++t) {
            int np = v[t].q1 - v[t].q0;
            while (lg[t] < lgmax && np > (8 << lg[t])) ++lg[t];
            total += 1 << lg[t];
        }
        // narrow phase: spread every task over more lanes while one round of the team can hold them
        while (total * 2 <= team_lanes) {
            bool any = false;
            total = 0;
            for (size_t t = 0; t < v.size(); ++t) {
                if (lg[t] < lgmax) { ++lg[t]; any = true; }
                total += 1 << lg[t];
            }
            if (!any) break;
        }
        int s0 = (int)out.ftask.size() / 4, mp = 0, mlg = 0;
        for (size_t t = 0; t < v.size(); ++t) {
            const Tk& k = v[t];
            for (int lane = 0; lane < (1 << lg[t]); ++lane) {
                out.ftask.push_back(k.sw ? (int)((unsigned)k.tgt | ((unsigned)lg[t] << 26) | (lane == 0 ? (1u << 29) : 0u) | (1u << 31))
                                         : (k.tgt | (lg[t] << 26) | (lane == 0 ? (1 << 29) : 0) | (S.hasK[k.tgt] ? (1 << 30) : 0)));
                out.ftask.push_back(k.q0 + lane); out.ftask.push_back(k.q1); out.ftask.push_back(k.aux);
            }
            mp = std::max(mp, k.q1 - k.q0);
            mlg = std::max(mlg, lg[t]);
        }
        out.ftasks += (int)v.size();
        // phase = (first slot, end slot, max pairs of a task | max lg << 24, kind)
        out.fphase.push_back(s0); out.fphase.push_back((int)out.ftask.size() / 4); out.fphase.push_back(mp | (mlg << 24)); out.fphase.push_back(kind);
        v.clear();
    };
    std::vector<Tk> v;
    for (int l = 0; l < S.nlev; ++l) {
        for (int j = S.lev_ptr[l]; j < S.lev_ptr[l + 1]; ++j) v.push_back(Tk{S.Lp[j], S.fp_ptr[S.Lp[j]], S.fp_ptr[S.Lp[j] + 1], j});
        emit(v, 0);
        for (int j = S.lev_ptr[l]; j < S.lev_ptr[l + 1]; ++j) {
            for (int e = S.Lp[j] + 1; e < S.Lp[j + 1]; ++e) v.push_back(Tk{e, S.fp_ptr[e], S.fp_ptr[e + 1], j});
            if (fuse_fwd) { Tk t{j, S.sw_off + S.Rp[j], S.sw_off + S.Rp[j + 1], j}; t.sw = true; v.push_back(t); }
        }
        if (!v.empty()) emit(v, 1);
    }
    if (S.T > 0) {
        // packed position (row-major lower triangle of the T x T tail) of every tail entry
        for (int j = n0; j < n; ++j)
            for (int e = S.Lp[j]; e < S.Lp[j + 1]; ++e) {
                int r = S.Li[e] - n0, c = j - n0;
                v.push_back(Tk{e, S.fp_ptr[e], S.fp_ptr[e + 1], r * (r + 1) / 2 + c});
            }
        if (fuse_fwd)  // right-hand side of the tail: yw[j] -= sum over the sparse columns
            for (int j = n0; j < n; ++j)
                if (S.Rmid[j] > S.Rp[j]) { Tk t{j, S.sw_off + S.Rp[j], S.sw_off + S.Rmid[j], j}; t.sw = true; v.push_back(t); }
        emit(v, 2);
    }
}
