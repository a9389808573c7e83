// TEST INFRASTRUCTURE ONLY (not part of libsqpqp.so, never on the product path).
// Executes the index programs produced by csrc/symbolic.hpp on the host, in the order the
// device kernels execute them (csrc/chol.cuh: assembly, level-scheduled sparse columns, Schur
// complement of the dense tail, dense packed Cholesky of the tail, forward / tail / backward
// sweeps), so that tests/test_symbolic.py can validate the symbolic analysis against a dense
// Cholesky on a machine without a GPU.
#include <cmath>
#include <cstdint>
#include <vector>
#include "../../sqpsolver.jl_b200/csrc/symbolic.hpp"

static inline int tri(int r) { return r * (r + 1) / 2; }

extern "C" int symcheck_solve3(int n, int m, const int* Jrp, const int* Jcol, const double* Jv, const int* Prp, const int* Pcol,
                               const double* Pv, const double* d, const double* w, const double* rhs, double* x,
                               int64_t* stats /* nnzL, sparse levels, flops, assembly terms, tail, tree levels */, int tail_max, int fuse) {
    Symbolic S = symbolic_analyze(n, m, Jrp, Jrp + 1, Jcol, Prp, Pcol, 512, tail_max, fuse != 0);
    if (!S.ok) return -1;
    if ((fuse != 0) != S.fused_fwd) return -12;
    const int n0 = S.n0, T = S.T;
    std::vector<double> L(S.nnzL, std::nan("")), D((size_t)tri(T) + T, 0.0), dinv(n), wJ(S.jrow.size());
    if (S.nnzL >= (1 << 26)) return -9;  // slot encoding: 26 bits of entry id
    // assembly (chol_assemble): wJ, then the diagonal and the sourced sub-diagonal entries; pure fill stays unset (NaN here:
    // a factor task that read it although has_K = 0 would poison the result)
    for (size_t a = 0; a < S.jrow.size(); ++a) wJ[a] = S.jrow[a] >= 0 ? w[S.jrow[a]] * Jv[a] : 0.0;
    for (size_t t = 0; t < S.aslot_d.size();) {  // assembly slots: a task = 2^lg consecutive slots, leader first
        const int* tk = &S.aslot[4 * t];
        const int e = tk[0] & 0x3ffffff, Ln = 1 << ((tk[0] >> 26) & 7);
        if (!((tk[0] >> 29) & 1) || t % Ln != 0) return -4;
        double v = (Pv && tk[3] >= 0) ? Pv[tk[3]] : 0.0;
        if (S.aslot_d[t] >= 0) v += d[S.aslot_d[t]];
        for (int lane = 0; lane < Ln; ++lane) {
            const int* sl = &S.aslot[4 * (t + lane)];
            if ((sl[0] & 0x3ffffff) != e || sl[1] != tk[1] + lane) return -4;
            for (int q = sl[1]; q < sl[2]; q += Ln) v += wJ[S.as_ab[2 * (size_t)q]] * Jv[S.as_ab[2 * (size_t)q + 1]];
        }
        L[e] = v;
        t += Ln;
    }
    // factorisation phases (chol_factor).  done[e] = phase that produced entry e: a task may only read entries of
    // EARLIER phases (the device runs the tasks of one phase concurrently).
    std::vector<int> done(S.nnzL, 1 << 30), ddone(n, 1 << 30), ydone(n, fuse ? (1 << 30) : -1);
    std::vector<double> y(n);
    for (int k = 0; k < n; ++k) y[k] = rhs[S.perm[k]];
    for (size_t p = 0; p < S.fphase.size() / 4; ++p) {
        const int* ph = &S.fphase[4 * p];
        for (int t = ph[0]; t < ph[1];) {
            const int* tk = &S.ftask[4 * (size_t)t];
            const int e = tk[0] & 0x3ffffff, Ln = 1 << ((tk[0] >> 26) & 7);
            const bool sw = tk[0] < 0;                          // forward-sweep slot of a fused program
            if (sw && !fuse) return -13;
            if (!((tk[0] >> 29) & 1)) return -6;                // a task starts with its leader slot
            if ((t - ph[0]) % Ln != 0) return -7;               // lane groups aligned to their size
            double acc = 0.0;
            int npairs = 0;
            for (int lane = 0; lane < Ln; ++lane) {             // the lanes of the task, each with stride Ln
                const int* sl = &S.ftask[4 * (size_t)(t + lane)];
                if ((sl[0] & 0x3ffffff) != e || sl[1] != tk[1] + lane || (lane > 0 && ((sl[0] >> 29) & 1)) || ((sl[0] < 0) != sw)) return -8;
                for (int q = sl[1]; q < sl[2]; q += Ln, ++npairs) {
                    const int ea = S.fp_ab[2 * (size_t)q], eb = S.fp_ab[2 * (size_t)q + 1];
                    if (sw) {
                        if (done[ea] >= (int)p || ydone[eb] >= (int)p) return -10;
                        acc += L[ea] * y[eb];
                    } else {
                        if (done[ea] >= (int)p || done[eb] >= (int)p) return -10;  // intra-phase dependency
                        acc += L[ea] * L[eb];
                    }
                }
            }
            if (sw) {  // row e of the forward sweep (levels) or the right-hand side of the tail (Schur phase)
                if (npairs != tk[2] - tk[1]) return -5;
                if (ph[3] == 1) { if (ddone[e] >= (int)p) return -11; y[e] = (y[e] - acc) * dinv[e]; }
                else if (ph[3] == 2) y[e] -= acc;
                else return -14;
                ydone[e] = (int)p;
                t += Ln;
                continue;
            }
            if (npairs != tk[2] - tk[1] || npairs > (ph[2] & 0xffffff)) return -5;  // ph[2] = max pairs | max lg << 24
            double v = ((tk[0] >> 30) & 1 ? L[e] : 0.0) - acc;
            if (ph[3] == 0) {
                if (!(v > 0.0)) return -2;
                double inv = 1.0 / std::sqrt(v);
                L[e] = v * inv;
                dinv[tk[3]] = inv;
                ddone[tk[3]] = (int)p;
            } else if (ph[3] == 1) {
                if (ddone[tk[3]] >= (int)p) return -11;
                L[e] = v * dinv[tk[3]];
            } else {
                D[tk[3]] = v;
            }
            done[e] = (int)p;
            t += Ln;
        }
    }
    if (S.nlev > 0 && S.lev_ptr[S.nlev] != n0) return -3;
    if (T > 0) {  // left-looking dense Cholesky of the packed tail (dense_factor)
        for (int j = 0; j < T; ++j) {
            double dj = 0.0;
            for (int k = 0; k < j; ++k) dj += D[tri(j) + k] * D[tri(j) + k];
            double pv = D[tri(j) + j] - dj;
            if (!(pv > 0.0)) return -2;
            double inv = 1.0 / std::sqrt(pv);
            dinv[n0 + j] = inv;
            for (int i = j + 1; i < T; ++i) {
                double di = 0.0;
                for (int k = 0; k < j; ++k) di += D[tri(i) + k] * D[tri(j) + k];
                D[tri(i) + j] = (D[tri(i) + j] - di) * inv;
            }
        }
    }
    for (int j = 0; j < n0 && !fuse; ++j) {
        double acc = 0.0;
        for (int q = S.Rp[j]; q < S.Rp[j + 1]; ++q) acc += L[S.Rci[2 * (size_t)q]] * y[S.Rci[2 * (size_t)q + 1]];
        y[j] = (y[j] - acc) * dinv[j];
    }
    if (fuse)
        for (int j = 0; j < n0; ++j) if (ydone[j] > (1 << 29)) return -15;  // every row of the sparse levels was swept
    if (T > 0) {
        for (int j = n0; j < n && !fuse; ++j) {
            double acc = 0.0;
            for (int q = S.Rp[j]; q < S.Rmid[j]; ++q) acc += L[S.Rci[2 * (size_t)q]] * y[S.Rci[2 * (size_t)q + 1]];
            y[j] -= acc;
        }
        double* t = &y[n0];
        for (int j = 0; j < T; ++j) {
            t[j] *= dinv[n0 + j];
            for (int i = j + 1; i < T; ++i) t[i] -= D[tri(i) + j] * t[j];
        }
        for (int j = T - 1; j >= 0; --j) {
            t[j] *= dinv[n0 + j];
            for (int k = 0; k < j; ++k) t[k] -= D[tri(j) + k] * t[j];
        }
    }
    for (int j = n0 - 1; j >= 0; --j) {
        double acc = 0.0;
        for (int p = S.Lp[j] + 1; p < S.Lp[j + 1]; ++p) acc += L[p] * y[S.Li[p]];
        y[j] = (y[j] - acc) * dinv[j];
    }
    for (int k = 0; k < n; ++k) x[S.perm[k]] = y[k];
    if (stats) {
        stats[0] = S.nnzL; stats[1] = S.nlev; stats[2] = S.flops; stats[3] = S.as_terms;
        stats[4] = S.T; stats[5] = S.nlev_total;
    }
    return 0;
}

extern "C" int symcheck_solve2(int n, int m, const int* Jrp, const int* Jcol, const double* Jv, const int* Prp, const int* Pcol,
                               const double* Pv, const double* d, const double* w, const double* rhs, double* x, int64_t* stats,
                               int tail_max) {
    return symcheck_solve3(n, m, Jrp, Jcol, Jv, Prp, Pcol, Pv, d, w, rhs, x, stats, tail_max, 0);
}

extern "C" int symcheck_solve(int n, int m, const int* Jrp, const int* Jcol, const double* Jv, const int* Prp, const int* Pcol,
                              const double* Pv, const double* d, const double* w, const double* rhs, double* x, int64_t* stats) {
    int64_t st[6];
    int rc = symcheck_solve2(n, m, Jrp, Jcol, Jv, Prp, Pcol, Pv, d, w, rhs, x, stats ? st : nullptr, 0);
    if (stats && rc == 0) for (int k = 0; k < 4; ++k) stats[k] = st[k];
    return rc;
}

// Executes the RING program (symbolic.hpp: build_ring_program) the way the resident CTA team does (chol.cuh: ring_run_segment):
// chunk by chunk, every slot sums its own pairs (k-major words; factor / sweep chunks run all kmax pairs, the padding multiplies
// the zero entry L[nL]), the lanes of a task are reduced, the leader finalises by the slot's kind.  Within a chunk every slot
// reads the state BEFORE the chunk (the device runs the slots of a chunk concurrently), which the deferred write list
// reproduces.  Segment 0 factorises AND forward-solves; a second solve with the same factor then goes through segment 1.
extern "C" int symcheck_ring(int n, int m, const int* Jrp, const int* Jcol, const double* Jv, const int* Prp, const int* Pcol,
                             const double* Pv, const double* d, const double* w, const double* rhs, double* x, int tail_max,
                             int ns_max, int kcap, int stage_bytes, int stages, int64_t* stats /* chunks, words, stage_words, nL */) {
    Symbolic S = symbolic_analyze(n, m, Jrp, Jrp + 1, Jcol, Prp, Pcol, 512, tail_max);
    if (!S.ok) return -1;
    RingProg R;
    build_ring_program(S, Prp != nullptr, ns_max, kcap, stage_bytes, stages, R);
    if (!R.ok) return -20;
    const int n0 = S.n0, T = S.T;
    std::vector<double> L((size_t)R.nL + 1, std::nan("")), D((size_t)tri(T) + T, 0.0), dinv(n, std::nan("")), wJ(S.jrow.size()), y(n);
    L[R.nL] = 0.0;
    for (size_t a = 0; a < S.jrow.size(); ++a) wJ[a] = S.jrow[a] >= 0 ? w[S.jrow[a]] * Jv[a] : 0.0;
    struct Wr { int arr, idx; double v; };  // arr: 0 L, 1 D, 2 dinv, 3 y
    auto run_segment = [&](int seg) -> int {
        for (int c = R.seg_first[seg]; c < R.seg_first[seg] + R.seg_count[seg]; ++c) {
            const int* W = &R.words[R.chunk_off[c]];
            const int npad = W[0], kmax = W[1], is_asm = W[2];
            if (R.chunk_len[c] != 8 + 2 * npad + kmax * npad || R.chunk_len[c] * 4 > stage_bytes || (npad & 31) || (kmax & 1)) return -21;
            if (W[3] != (c + 1 == R.seg_first[seg] + R.seg_count[seg])) return -22;
            if (c + stages < R.seg_first[seg] + R.seg_count[seg]) { if (W[4] != R.chunk_off[c + stages] || W[5] != 4 * R.chunk_len[c + stages]) return -23; }
            else if (W[4] != -1) return -23;
            const int* sl = W + 8;
            const int* pw = W + 8 + 2 * npad;
            std::vector<Wr> wr;
            for (int s = W[6]; s < npad; ++s)  // idle slots of a factor / sweep chunk must be harmless: zero entry only
                for (int k = 0; k < kmax && !is_asm; ++k)
                    if (((unsigned)pw[k * npad + s] & 0xffffu) != (unsigned)R.nL) return -29;
            for (int s = 0; s < W[6];) {
                const int w0 = sl[2 * s], aux = sl[2 * s + 1];
                const int tgt = w0 & 0xffff, Ln = 1 << ((w0 >> 16) & 7), hasK = (w0 >> 20) & 1, kind = (w0 >> 28) & 7;
                if ((kind >= 5) != (is_asm != 0)) return -31;
                if (!((w0 >> 19) & 1) || s % Ln != 0) return -24;  // leader first, group aligned to its size
                double acc = 0.0;
                for (int lane = 0; lane < Ln; ++lane) {
                    const int v0 = sl[2 * (s + lane)];
                    if ((v0 & 0xffff) != tgt || (lane > 0 && ((v0 >> 19) & 1)) || ((v0 >> 28) & 7) != kind) return -25;
                    const int ks = (v0 >> 21) & 0x7f;
                    if (ks > kmax) return -26;
                    for (int k = 0; k < (is_asm ? ks : kmax); ++k) {
                        const unsigned p = (unsigned)pw[k * npad + s + lane];
                        const int a = p & 0xffff, b = p >> 16;
                        double va, vb;
                        if (kind <= 2) { if (a > R.nL || b > R.nL) return -27; va = L[a]; vb = L[b]; }
                        else if (kind <= 4) { if (a > R.nL || b >= n) return -27; va = L[a]; vb = (a == R.nL) ? 0.0 : y[b]; }
                        else { va = wJ[a]; vb = Jv[b]; }
                        if (va != va || vb != vb) return -28;  // read of a value no earlier chunk produced
                        acc += va * vb;
                    }
                }
                if (kind == 0) {
                    double v = (hasK ? L[tgt] : 0.0) - acc;
                    if (!(v > 0.0)) return -2;
                    const double inv = 1.0 / std::sqrt(v);
                    wr.push_back({0, tgt, v * inv}); wr.push_back({2, aux, inv});
                } else if (kind == 1) {
                    wr.push_back({0, tgt, ((hasK ? L[tgt] : 0.0) - acc) * dinv[aux]});
                } else if (kind == 2) {
                    wr.push_back({1, aux, D[aux] - acc});
                } else if (kind == 3) {
                    wr.push_back({3, tgt, (y[tgt] - acc) * dinv[tgt]});
                } else if (kind == 4) {
                    wr.push_back({3, tgt, y[tgt] - acc});
                } else {
                    const int h = (aux & 0xffff) - 1, dd = (int)((unsigned)aux >> 16) - 1;
                    double v = acc;
                    if (Pv && h >= 0) v += Pv[h];
                    if (dd >= 0) v += d[dd];
                    wr.push_back({kind == 5 ? 0 : 1, tgt, v});
                }
                s += Ln;
            }
            for (const Wr& q : wr) {
                if (q.arr == 0) L[q.idx] = q.v; else if (q.arr == 1) D[q.idx] = q.v; else if (q.arr == 2) dinv[q.idx] = q.v; else y[q.idx] = q.v;
            }
        }
        return 0;
    };
    auto tail_solve = [&]() {
        if (T <= 0) return;
        double* t = &y[n0];
        for (int j = 0; j < T; ++j) {
            t[j] *= dinv[n0 + j];
            for (int i = j + 1; i < T; ++i) t[i] -= D[tri(i) + j] * t[j];
        }
        for (int j = T - 1; j >= 0; --j) {
            t[j] *= dinv[n0 + j];
            for (int k = 0; k < j; ++k) t[k] -= D[tri(j) + k] * t[j];
        }
    };
    for (int k = 0; k < n; ++k) y[k] = rhs[S.perm[k]];
    int rc = run_segment(0);  // assembly, factorisation, forward sweep
    if (rc) return rc;
    if (T > 0) {
        for (int j = 0; j < T; ++j) {
            double dj = 0.0;
            for (int k = 0; k < j; ++k) dj += D[tri(j) + k] * D[tri(j) + k];
            double pv = D[tri(j) + j] - dj;
            if (!(pv > 0.0)) return -2;
            double inv = 1.0 / std::sqrt(pv);
            dinv[n0 + j] = inv;
            for (int i = j + 1; i < T; ++i) {
                double di = 0.0;
                for (int k = 0; k < j; ++k) di += D[tri(i) + k] * D[tri(j) + k];
                D[tri(i) + j] = (D[tri(i) + j] - di) * inv;
            }
        }
    }
    tail_solve();
    if ((rc = run_segment(2))) return rc;
    for (int k = 0; k < n; ++k) x[S.perm[k]] = y[k];
    // a second solve with the finished factor: forward sweep on its own (segment 1), tail, backward sweep
    std::vector<double> x2(n);
    for (int k = 0; k < n; ++k) y[k] = rhs[S.perm[k]];
    if ((rc = run_segment(1))) return rc;
    tail_solve();
    if ((rc = run_segment(2))) return rc;
    double mx = 0.0, df = 0.0;
    for (int k = 0; k < n; ++k) { x2[S.perm[k]] = y[k]; }
    for (int k = 0; k < n; ++k) { mx = std::fmax(mx, std::fabs(x[k])); df = std::fmax(df, std::fabs(x[k] - x2[k])); }
    if (!(df <= 1e-9 * std::fmax(1.0, mx))) return -30;
    if (stats) { stats[0] = R.nchunks; stats[1] = (int64_t)R.words.size(); stats[2] = R.stage_words; stats[3] = R.nL; }
    return 0;
}
