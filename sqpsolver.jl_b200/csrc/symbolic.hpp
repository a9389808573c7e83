// symbolic.hpp -- host-side symbolic analysis of the condensed Newton/KKT matrix
//
//      K = P + diag(d) + J' diag(w) J            (n x n, symmetric positive definite)
//
// shared by every instance of a batch (one sparsity pattern) and by every iteration of a
// solve (only d, w and the values of P, J change).  Done once at setup, in plain C++:
//   1. pattern of K = pattern(P) U pattern(J'J) U I
//   2. fill-reducing ordering: minimum degree on the elimination graph (exact external degree)
//   3. symbolic Cholesky: column structure of L (by-product of 2.), elimination-tree levels
//   4. the three index programs the device kernels execute:
//        assembly : K_e  = P[h_idx] + d[diag] + sum_t w[row_t] * Jv[a_t] * Jv[b_t]
//        factor   : L_e  = (K_e - sum_t L[p_t] * L[q_t]) / L_jj       (level by level)
//        solve    : level-scheduled forward (rows of L) and backward (columns of L) sweeps
// No numerical work happens here.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

struct Symbolic {
    int n = 0, nnzL = 0, nlev = 0;
    std::vector<int> perm, iperm;        // perm[k] = original index of pivot k
    std::vector<int> Lp, Li;             // CSC of L (permuted indices), diagonal first in each column
    std::vector<int> Rp, Rc, Ri;         // CSR of strictly-lower L: row ptr, column, index into the CSC value array
    std::vector<int> lev_ptr, lev_cols;  // columns grouped by elimination-tree level (leaves first)
    // factorisation program: entries ordered by (level, diagonal-before-offdiagonal)
    std::vector<int> fd_ptr, fo_ptr;     // per level: range in f_ent of diagonal / off-diagonal entries
    std::vector<int> f_ent;              // entry ids (index into L values)
    std::vector<int> fp_ptr;             // per entry id: range in fp_a/fp_b
    std::vector<int> fp_a, fp_b;         // pairs of L value indices to multiply-subtract
    std::vector<int> ent_diag;           // per entry id: value index of the diagonal of its column
    // assembly program
    std::vector<int> as_ptr;             // per entry id: range in as_a/as_b/as_r
    std::vector<int> as_a, as_b, as_r;   // Jv index a, Jv index b, row (weight index)
    std::vector<int> as_h;               // per entry id: index into P values or -1
    std::vector<int> as_d;               // per entry id: original column for d[] (diagonal entries) or -1
    int64_t flops = 0;
    bool ok = false;                     // false: a row of J is too long for the clique expansion
};

// J: m x ncols CSR (rb/re per row, so a prefix of each row can be used), P: symmetric-full CSR or null
inline Symbolic symbolic_analyze(int n, int m, const int* Jrb, const int* Jre, const int* Jcol, const int* Prp,
                                 const int* Pcol, int max_row_len = 512) {
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

    // ---- 2./3. minimum degree elimination; column structures fall out ----------------------
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
    // elimination tree + levels
    std::vector<int> parent(n, -1), level(n, 0);
    for (int k = 0; k < n; ++k)
        if (S.Lp[k + 1] - S.Lp[k] > 1) parent[k] = S.Li[S.Lp[k] + 1];
    for (int k = 0; k < n; ++k)
        if (parent[k] >= 0) level[parent[k]] = std::max(level[parent[k]], level[k] + 1);
    S.nlev = 0;
    for (int k = 0; k < n; ++k) S.nlev = std::max(S.nlev, level[k] + 1);
    S.lev_ptr.assign(S.nlev + 1, 0);
    for (int k = 0; k < n; ++k) S.lev_ptr[level[k] + 1]++;
    for (int l = 0; l < S.nlev; ++l) S.lev_ptr[l + 1] += S.lev_ptr[l];
    S.lev_cols.resize(n);
    {
        std::vector<int> cur(S.lev_ptr.begin(), S.lev_ptr.end() - 1);
        for (int k = 0; k < n; ++k) S.lev_cols[cur[level[k]]++] = k;
    }
    // CSR of strictly-lower L (row i: columns k < i)
    S.Rp.assign(n + 1, 0);
    for (int k = 0; k < n; ++k)
        for (int p = S.Lp[k] + 1; p < S.Lp[k + 1]; ++p) S.Rp[S.Li[p] + 1]++;
    for (int i = 0; i < n; ++i) S.Rp[i + 1] += S.Rp[i];
    S.Rc.resize(S.Rp[n]);
    S.Ri.resize(S.Rp[n]);
    {
        std::vector<int> cur(S.Rp.begin(), S.Rp.end() - 1);
        for (int k = 0; k < n; ++k)
            for (int p = S.Lp[k] + 1; p < S.Lp[k + 1]; ++p) {
                int i = S.Li[p];
                S.Rc[cur[i]] = k;
                S.Ri[cur[i]++] = p;
            }
    }
    // ---- 4a. factorisation program ------------------------------------------------------------
    // entry (i,j), j<=i:  sum over k<j with L_ik != 0 and L_jk != 0 -> intersect rows i and j of the CSR
    S.fp_ptr.assign(S.nnzL + 1, 0);
    S.ent_diag.resize(S.nnzL);
    for (int pass = 0; pass < 2; ++pass) {
        int64_t total = 0;
        for (int j = 0; j < n; ++j)
            for (int p = S.Lp[j]; p < S.Lp[j + 1]; ++p) {
                int i = S.Li[p];
                S.ent_diag[p] = S.Lp[j];
                int a = S.Rp[i], ae = S.Rp[i + 1], b = S.Rp[j], be = S.Rp[j + 1];
                int cnt = 0;
                while (a < ae && b < be) {
                    int ca = S.Rc[a], cb = S.Rc[b];
                    if (ca >= j || cb >= j) break;
                    if (ca == cb) {
                        if (pass) { S.fp_a[S.fp_ptr[p] + cnt] = S.Ri[a]; S.fp_b[S.fp_ptr[p] + cnt] = S.Ri[b]; }
                        ++cnt; ++a; ++b;
                    } else if (ca < cb) ++a; else ++b;
                }
                if (!pass) S.fp_ptr[p + 1] = cnt;
                total += cnt;
            }
        if (!pass) {
            for (int p = 0; p < S.nnzL; ++p) S.fp_ptr[p + 1] += S.fp_ptr[p];
            S.fp_a.resize(S.fp_ptr[S.nnzL]);
            S.fp_b.resize(S.fp_ptr[S.nnzL]);
            S.flops = 2 * total;
        }
    }
    S.fd_ptr.assign(S.nlev + 1, 0);
    S.fo_ptr.assign(S.nlev + 1, 0);
    S.f_ent.clear();
    // layout of f_ent: for each level: [diag entries][offdiag entries]
    for (int l = 0; l < S.nlev; ++l) {
        S.fd_ptr[l] = (int)S.f_ent.size();
        for (int t = S.lev_ptr[l]; t < S.lev_ptr[l + 1]; ++t) S.f_ent.push_back(S.Lp[S.lev_cols[t]]);
        S.fo_ptr[l] = (int)S.f_ent.size();
        for (int t = S.lev_ptr[l]; t < S.lev_ptr[l + 1]; ++t) {
            int j = S.lev_cols[t];
            for (int p = S.Lp[j] + 1; p < S.Lp[j + 1]; ++p) S.f_ent.push_back(p);
        }
    }
    S.fd_ptr[S.nlev] = S.fo_ptr[S.nlev] = (int)S.f_ent.size();

    // ---- 4b. assembly program --------------------------------------------------------------------
    // map (i,j) permuted, i>=j -> entry id
    S.as_h.assign(S.nnzL, -1);
    S.as_d.assign(S.nnzL, -1);
    auto find_entry = [&](int a, int b) {  // original indices
        int i = S.iperm[a], j = S.iperm[b];
        if (i < j) std::swap(i, j);
        const int* lo = &S.Li[S.Lp[j]] + 1;
        const int* hi = &S.Li[0] + S.Lp[j + 1];
        if (i == j) return S.Lp[j];
        const int* it = std::lower_bound(lo, hi, i);
        return (int)(it - &S.Li[0]);
    };
    for (int k = 0; k < n; ++k) S.as_d[S.Lp[k]] = S.perm[k];
    if (Prp)
        for (int a = 0; a < n; ++a)
            for (int k = Prp[a]; k < Prp[a + 1]; ++k) {
                int b = Pcol[k];
                if (S.iperm[a] >= S.iperm[b]) S.as_h[find_entry(a, b)] = k;  // lower triangle in permuted order
            }
    std::vector<int> cnt(S.nnzL + 1, 0);
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
                        int pos = S.as_ptr[e] + cnt[e]++;
                        S.as_a[pos] = a; S.as_b[pos] = b; S.as_r[pos] = r;
                    }
                }
        if (!pass) {
            S.as_ptr.assign(S.nnzL + 1, 0);
            for (int e = 0; e < S.nnzL; ++e) S.as_ptr[e + 1] = S.as_ptr[e] + cnt[e + 1];
            S.as_a.resize(S.as_ptr[S.nnzL]);
            S.as_b.resize(S.as_ptr[S.nnzL]);
            S.as_r.resize(S.as_ptr[S.nnzL]);
            std::fill(cnt.begin(), cnt.end(), 0);
        }
    }
    S.ok = true;
    return S;
}
