// acopf.cuh -- device-side evaluator of the polar AC optimal-power-flow NLP for the batched workload
// (SURVEY.md section 8f, rank 1): f, grad f, g, the Jacobian COO values and the Lagrangian-Hessian COO values,
// written straight into the engine's device input buffers in exactly the COO order the host handed to
// sqpqp_setup_nlp, so that the ordered scatter (pattern.cuh, sqp.jl:92-117) consumes them unchanged.
//
// It replaces, for this NLP family, what the reference obtains from the MOI NLPEvaluator callbacks in
// eval_functions! (sqp.jl:86-104: eval_f, eval_grad_f, eval_g, eval_jac_g, eval_h with sigma = 1 and mu = lambda)
// and the per-iteration upload of their results.  Formulation and COO ordering: PowerModels ACPPowerModel +
// build_opf as the reference's test/opf.jl:5-9 builds it, rows ordered as MOI_wrapper.jl:759-766 (see
// sqpsolver.jl_b200/nlp/acopf.py, which is the host twin of this file and the parity reference of its test):
//
//   variables  va[nb] vm[nb] pg[ng] qg[ng] p[2 nl] q[2 nl]            (arcs: from [nl], then to [nl])
//   rows       angle-diff <= [nl], angle-diff >= [nl], ref angle, thermal from/to interleaved [2 nl],
//              balance P,Q interleaved [2 nb], Ohm p_fr,q_fr,p_to,q_to per branch [4 nl]
//   J COO      affine [4 nl + 1], thermal [4 nl], balance entries [nbal], Ohm 5 per row [20 nl]
//   H COO      objective [ng], thermal [4 nl], shunt P,Q per shunt bus [2 nsh], Ohm 9 per row [36 nl]
//
// One CTA per instance; every output element is written by exactly one thread (no atomics).
#pragma once
#include "team.cuh"

struct AcopfDev {
    int nb, ng, nl, ref_bus, nbal, nsh;
    const int *f_bus, *t_bus, *gen_bus;
    const double *oa, *oc, *os;            // [nl][4] Ohm-row coefficients
    const double *cost2, *cost1, *cost0;   // [ng]
    const double *gs, *bs;                 // [nb]
    const int *bal_ptr, *bal_col, *bal_kind;  // balance rows as CSR over the J COO entries: kind 0 constant, 1 P shunt, 2 Q shunt
    const double* bal_const;
    const int* sh_bus;                     // [nsh]
};

struct AcopfArgs {
    const double *x, *lam;   // [batch][n], [batch][m]
    const int* mask;         // nullable: evaluate only instances with mask != 0
    double *dE, *hval, *df, *E, *f;
    int nnzJ, nnzH;
    int fonly;               // 1: function values only (f and E at a trial point: compute_phi, sqp.jl:170-183); dE / hval / df untouched
};

__global__ void __launch_bounds__(256) k_acopf_eval(AcopfDev A, AcopfArgs G, int n, int m, int batch) {
    __shared__ double sh[2 * SQPQP_MAX_RED * 32];
    const int nb = A.nb, ng = A.ng, nl = A.nl;
    const int o_va = 0, o_vm = nb, o_pg = 2 * nb, o_qg = 2 * nb + ng, o_p = 2 * nb + 2 * ng, o_q = o_p + 2 * nl;
    const int r_angU = 0, r_angL = nl, r_ref = 2 * nl, r_th = 2 * nl + 1, r_bal = r_th + 2 * nl, r_ohm = r_bal + 2 * nb;
    const int j_th = 4 * nl + 1, j_bal = j_th + 4 * nl, j_ohm = j_bal + A.nbal;
    const int h_th = ng, h_sh = ng + 4 * nl, h_ohm = h_sh + 2 * A.nsh;
    for (int inst = blockIdx.x; inst < batch; inst += gridDim.x) {
        if (G.mask && !G.mask[inst]) continue;
        CtaTeam T(sh);
        const double* x = G.x + (size_t)inst * n;
        const double* lam = G.lam + (size_t)inst * m;
        double* E = G.E + (size_t)inst * m;
        const int tid = threadIdx.x, nt = blockDim.x;
        if (G.fonly) {  // the same expressions as below, values only (bit-identical E and f)
            double fo[1] = {0.0};
            for (int k = tid; k < ng; k += nt) {
                const double pg = x[o_pg + k], c2 = A.cost2[k], c1 = A.cost1[k];
                fo[0] += c2 * pg * pg + c1 * pg + A.cost0[k];
            }
            for (int l = tid; l < nl; l += nt) {
                const int fb = A.f_bus[l], tb = A.t_bus[l];
                const double dth = x[o_va + fb] - x[o_va + tb];
                E[r_angU + l] = dth;
                E[r_angL + l] = dth;
                const double pf = x[o_p + l], qf = x[o_q + l], pt = x[o_p + nl + l], qt = x[o_q + nl + l];
                E[r_th + 2 * l] = pf * pf + qf * qf;
                E[r_th + 2 * l + 1] = pt * pt + qt * qt;
            }
            if (tid == 0) E[r_ref] = x[o_va + A.ref_bus];
            for (int r = tid; r < 2 * nb; r += nt) {
                const int bus = r >> 1;
                double acc = 0.0;
                for (int k = A.bal_ptr[r]; k < A.bal_ptr[r + 1]; ++k) {
                    const int kind = A.bal_kind[k];
                    const double xv = x[A.bal_col[k]];
                    if (kind == 0) acc += A.bal_const[k] * xv;
                    else if (kind == 1) acc += A.gs[bus] * xv * xv;
                    else acc -= A.bs[bus] * xv * xv;
                }
                E[r_bal + r] = acc;
            }
            for (int e = tid; e < 4 * nl; e += nt) {
                const int l = e >> 2, k = e & 3;
                const int fb = A.f_bus[l], tb = A.t_bus[l];
                const int bi = (k < 2) ? fb : tb, bj = (k < 2) ? tb : fb;
                const int var = (k == 0) ? o_p + l : (k == 1) ? o_q + l : (k == 2) ? o_p + nl + l : o_q + nl + l;
                const double vi = x[o_vm + bi], vj = x[o_vm + bj], th = x[o_va + bi] - x[o_va + bj];
                double sn, cs;
                sincos(th, &sn, &cs);
                const double C = A.oc[e] * cs + A.os[e] * sn;
                E[r_ohm + e] = x[var] - (A.oa[e] * vi * vi + vi * vj * C);
            }
            T.reduce<1, false>(fo);
            if (tid == 0) G.f[inst] = fo[0];
            __syncthreads();
            continue;
        }
        double* dE = G.dE + (size_t)inst * G.nnzJ;
        double* hv = G.hval + (size_t)inst * G.nnzH;
        double* df = G.df + (size_t)inst * n;
        // objective, gradient, its Hessian
        double fs[1] = {0.0};
        for (int j = tid; j < n; j += nt) df[j] = 0.0;
        __syncthreads();
        for (int k = tid; k < ng; k += nt) {
            const double pg = x[o_pg + k], c2 = A.cost2[k], c1 = A.cost1[k];
            fs[0] += c2 * pg * pg + c1 * pg + A.cost0[k];
            df[o_pg + k] = 2.0 * c2 * pg + c1;
            hv[k] = 2.0 * c2;
        }
        // angle-difference rows, reference angle
        for (int l = tid; l < nl; l += nt) {
            const int fb = A.f_bus[l], tb = A.t_bus[l];
            const double dth = x[o_va + fb] - x[o_va + tb];
            E[r_angU + l] = dth;
            E[r_angL + l] = dth;
            dE[2 * l] = 1.0; dE[2 * l + 1] = -1.0;
            dE[2 * nl + 2 * l] = 1.0; dE[2 * nl + 2 * l + 1] = -1.0;
        }
        if (tid == 0) { E[r_ref] = x[o_va + A.ref_bus]; dE[4 * nl] = 1.0; }
        // thermal rows: row 2l (from) = p_fr^2 + q_fr^2, row 2l+1 (to) = p_to^2 + q_to^2
        for (int l = tid; l < nl; l += nt) {
            const double pf = x[o_p + l], qf = x[o_q + l], pt = x[o_p + nl + l], qt = x[o_q + nl + l];
            E[r_th + 2 * l] = pf * pf + qf * qf;
            E[r_th + 2 * l + 1] = pt * pt + qt * qt;
            dE[j_th + 4 * l] = 2.0 * pf; dE[j_th + 4 * l + 1] = 2.0 * qf;
            dE[j_th + 4 * l + 2] = 2.0 * pt; dE[j_th + 4 * l + 3] = 2.0 * qt;
            const double l0 = 2.0 * lam[r_th + 2 * l], l1 = 2.0 * lam[r_th + 2 * l + 1];
            hv[h_th + 4 * l] = l0; hv[h_th + 4 * l + 1] = l0; hv[h_th + 4 * l + 2] = l1; hv[h_th + 4 * l + 3] = l1;
        }
        // balance rows: sum of arc flows - sum of generation (+ gs vm^2 / - bs vm^2)
        for (int r = tid; r < 2 * nb; r += nt) {
            const int bus = r >> 1;
            double acc = 0.0;
            for (int k = A.bal_ptr[r]; k < A.bal_ptr[r + 1]; ++k) {
                const int kind = A.bal_kind[k];
                const double xv = x[A.bal_col[k]];
                if (kind == 0) { acc += A.bal_const[k] * xv; dE[j_bal + k] = A.bal_const[k]; }
                else if (kind == 1) { acc += A.gs[bus] * xv * xv; dE[j_bal + k] = 2.0 * A.gs[bus] * xv; }
                else { acc -= A.bs[bus] * xv * xv; dE[j_bal + k] = -2.0 * A.bs[bus] * xv; }
            }
            E[r_bal + r] = acc;
        }
        for (int s = tid; s < A.nsh; s += nt) {
            const int bus = A.sh_bus[s];
            hv[h_sh + 2 * s] = 2.0 * A.gs[bus] * lam[r_bal + 2 * bus];
            hv[h_sh + 2 * s + 1] = -2.0 * A.bs[bus] * lam[r_bal + 2 * bus + 1];
        }
        // Ohm rows: var - [a vi^2 + vi vj (c cos(ti - tj) + s sin(ti - tj))], 4 rows per branch
        for (int e = tid; e < 4 * nl; e += nt) {
            const int l = e >> 2, k = e & 3;
            const int fb = A.f_bus[l], tb = A.t_bus[l];
            const int bi = (k < 2) ? fb : tb, bj = (k < 2) ? tb : fb;
            const int var = (k == 0) ? o_p + l : (k == 1) ? o_q + l : (k == 2) ? o_p + nl + l : o_q + nl + l;
            const double vi = x[o_vm + bi], vj = x[o_vm + bj], th = x[o_va + bi] - x[o_va + bj];
            double sn, cs;
            sincos(th, &sn, &cs);
            const double a = A.oa[e], c = A.oc[e], s = A.os[e];
            const double C = c * cs + s * sn, S = -c * sn + s * cs;
            E[r_ohm + e] = x[var] - (a * vi * vi + vi * vj * C);
            double* je = dE + j_ohm + 5 * e;
            const double dth = vi * vj * S;
            je[0] = 1.0; je[1] = -(2.0 * a * vi + vj * C); je[2] = -(vi * C); je[3] = -dth; je[4] = dth;
            const double lo = -lam[r_ohm + e], vvC = vi * vj * C;
            double* he = hv + h_ohm + 9 * e;
            he[0] = 2.0 * a * lo; he[1] = C * lo; he[2] = vj * S * lo; he[3] = -vj * S * lo; he[4] = vi * S * lo;
            he[5] = -vi * S * lo; he[6] = -vvC * lo; he[7] = vvC * lo; he[8] = -vvC * lo;
        }
        T.reduce<1, false>(fs);
        if (tid == 0) G.f[inst] = fs[0];
        __syncthreads();
    }
}
