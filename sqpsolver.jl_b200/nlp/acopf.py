"""Polar AC optimal power flow as an :class:`NLP`, with analytic J and H.

Formulation: PowerModels' ``ACPPowerModel`` + stock ``build_opf`` as used by the
reference's ``test/opf.jl:5-9`` (variables va, vm, pg, qg, p, q; theta-ref,
nodal balance, Ohm's law from/to, angle-difference and thermal-limit rows).
Row classes are ordered as ``MOI_wrapper.jl:759-766`` orders them:

    linear <=   angle difference upper           nbr rows
    linear >=   angle difference lower           nbr rows
    linear ==   reference angle                  1 row      -> num_linear = 2 nbr + 1
    quad   <=   thermal limit from / to          2 nbr rows (interleaved per branch)
    NLP         balance P,Q per bus              2 nbus rows
    NLP         Ohm p_fr,q_fr,p_to,q_to          4 nbr rows

so n = 2 nbus + 2 ngen + 4 nbr and m = 1 + 2 nbus + 8 nbr (case9: 60 / 91).
The Jacobian COO lists affine rows, then quadratic rows (each diagonal
quadratic term contributes one entry, MOI_wrapper.jl:913-927), then NLP rows
(MOI_wrapper.jl:930-945).  The Hessian COO lists the quadratic objective,
the quadratic-constraint terms, then the NLP block one triangle at a time with
one entry per (row, term) -- so slots shared by the four Ohm rows of a branch
appear four times, which exercises the ordered duplicate summation of
sqp.jl:92-103.

All callbacks accept leading batch dimensions on ``x`` / ``lam``; per-instance
loads can be supplied as ``pd[batch, nbus]`` (shared sparsity pattern).
"""
from __future__ import annotations

import numpy as np

from .base import NLP
from .networks import Network

INF = np.inf


class AcopfPolar(NLP):
    def subset(self, lo: int, hi: int) -> "AcopfPolar":
        """The NLP of instances lo..hi-1 of a batched (per-instance loads) problem; an unbatched problem is its own subset."""
        if self.pd.ndim == 1:
            return self
        return AcopfPolar(self.net, pd=self.pd[lo:hi], qd=self.qd[lo:hi])

    def __init__(self, net: Network, pd=None, qd=None):
        self.net = net
        self.name = f"acopf_{net.name}"
        nb, ng, nl = net.nbus, net.ngen, net.nbranch
        self.nb, self.ng, self.nl = nb, ng, nl
        self.pd = net.pd if pd is None else np.asarray(pd, float)
        self.qd = net.qd if qd is None else np.asarray(qd, float)
        # ---- variable layout -------------------------------------------------
        self.o_va, self.o_vm = 0, nb
        self.o_pg, self.o_qg = 2 * nb, 2 * nb + ng
        self.o_p = 2 * nb + 2 * ng  # p arcs: from [nl] then to [nl]
        self.o_q = self.o_p + 2 * nl
        self.n = n = 2 * nb + 2 * ng + 4 * nl
        x_L = np.full(n, -INF)
        x_U = np.full(n, INF)
        x_L[self.o_vm : self.o_vm + nb] = net.vmin
        x_U[self.o_vm : self.o_vm + nb] = net.vmax
        x_L[self.o_pg : self.o_pg + ng] = net.pmin
        x_U[self.o_pg : self.o_pg + ng] = net.pmax
        x_L[self.o_qg : self.o_qg + ng] = net.qmin
        x_U[self.o_qg : self.o_qg + ng] = net.qmax
        rate2 = np.concatenate([net.rate_a, net.rate_a])
        x_L[self.o_p : self.o_p + 2 * nl] = -rate2
        x_U[self.o_p : self.o_p + 2 * nl] = rate2
        x_L[self.o_q : self.o_q + 2 * nl] = -rate2
        x_U[self.o_q : self.o_q + 2 * nl] = rate2
        self.x_L, self.x_U = x_L, x_U
        # PowerModels start values: vm=1, everything else 0 (not clipped: the
        # MOI wrapper passes VariablePrimalStart through, MOI_wrapper.jl:1192-1199)
        x0 = np.zeros(n)
        x0[self.o_vm : self.o_vm + nb] = 1.0
        self.x0 = x0

        # ---- row layout ------------------------------------------------------
        self.r_angU, self.r_angL, self.r_ref = 0, nl, 2 * nl
        self.num_linear_constraints = 2 * nl + 1
        self.r_thermal = self.num_linear_constraints
        self.r_bal = self.r_thermal + 2 * nl
        self.r_ohm = self.r_bal + 2 * nb
        self.m = m = self.r_ohm + 4 * nl
        g_L = np.zeros(m)
        g_U = np.zeros(m)
        g_L[self.r_angU : self.r_angU + nl] = -INF
        g_U[self.r_angU : self.r_angU + nl] = net.angmax
        g_L[self.r_angL : self.r_angL + nl] = net.angmin
        g_U[self.r_angL : self.r_angL + nl] = INF
        g_L[self.r_thermal : self.r_thermal + 2 * nl] = -INF
        g_U[self.r_thermal : self.r_thermal + 2 * nl] = np.repeat(net.rate_a**2, 2)
        self.g_L, self.g_U = g_L, g_U
        self._set_balance_bounds()

        # ---- branch coefficients (PowerModels constraint_ohms_yt_from/to) ----
        f, t = net.f_bus, net.t_bus
        y = 1.0 / (net.br_r + 1j * net.br_x)
        g, b = y.real, y.imag
        tr = net.tap * np.cos(net.shift)
        ti = net.tap * np.sin(net.shift)
        tm2 = net.tap**2
        bfr = bto = net.br_b / 2.0
        # each Ohm row is  var - [a vi^2 + c vi vj cos(ti-tj) + s vi vj sin(ti-tj)]
        # rows per branch: 0 p_fr (i=f,j=t), 1 q_fr, 2 p_to (i=t,j=f), 3 q_to
        self.oa = np.stack([g / tm2, -(b + bfr) / tm2, g, -(b + bto)], 1)  # [nl,4]
        self.oc = np.stack(
            [(-g * tr + b * ti) / tm2, -(-b * tr - g * ti) / tm2, (-g * tr - b * ti) / tm2, -(-b * tr + g * ti) / tm2], 1
        )
        self.os = np.stack(
            [(-b * tr - g * ti) / tm2, (-g * tr + b * ti) / tm2, (-b * tr + g * ti) / tm2, (-g * tr - b * ti) / tm2], 1
        )
        ar = np.arange(nl)
        self.o_i = np.stack([f, f, t, t], 1)  # bus i per (branch,row)
        self.o_j = np.stack([t, t, f, f], 1)
        self.o_var = np.stack([self.o_p + ar, self.o_q + ar, self.o_p + nl + ar, self.o_q + nl + ar], 1)

        import scipy.sparse as sp

        ft = np.concatenate([f, t])
        self.M_arc = sp.csr_matrix((np.ones(2 * nl), (ft, np.arange(2 * nl))), shape=(nb, 2 * nl))
        self.M_gen = sp.csr_matrix((np.ones(ng), (net.gen_bus, np.arange(ng))), shape=(nb, ng))
        self._build_jacobian_structure()
        self._build_hessian_structure()

    # ------------------------------------------------------------------ bounds
    def _set_balance_bounds(self):
        nb = self.nb
        pd, qd = self.pd, self.qd
        if pd.ndim == 1:
            self.g_L[self.r_bal : self.r_bal + 2 * nb : 2] = -pd
            self.g_L[self.r_bal + 1 : self.r_bal + 2 * nb : 2] = -qd
            self.g_U[self.r_bal : self.r_bal + 2 * nb] = self.g_L[self.r_bal : self.r_bal + 2 * nb]
        else:  # batched loads: g_L/g_U become [batch, m]
            B = pd.shape[0]
            gl = np.broadcast_to(self.g_L, (B, self.m)).copy()
            gu = np.broadcast_to(self.g_U, (B, self.m)).copy()
            gl[:, self.r_bal : self.r_bal + 2 * nb : 2] = -pd
            gl[:, self.r_bal + 1 : self.r_bal + 2 * nb : 2] = -qd
            gu[:, self.r_bal : self.r_bal + 2 * nb] = gl[:, self.r_bal : self.r_bal + 2 * nb]
            self.g_L, self.g_U = gl, gu

    # --------------------------------------------------------------- structure
    def _build_jacobian_structure(self):
        net, nb, ng, nl = self.net, self.nb, self.ng, self.nl
        f, t = net.f_bus, net.t_bus
        ar = np.arange(nl)
        rows, cols = [], []
        # angle-difference <= and >= rows: (va_f, +1), (va_t, -1)
        for r0 in (self.r_angU, self.r_angL):
            rows.append(np.repeat(r0 + ar, 2))
            cols.append(np.stack([self.o_va + f, self.o_va + t], 1).ravel())
        # reference angle
        rows.append(np.array([self.r_ref]))
        cols.append(np.array([self.o_va + net.ref_bus]))
        self.j_n_affine = 4 * nl + 1
        self.j_affine_vals = np.concatenate([np.tile([1.0, -1.0], 2 * nl), [1.0]])
        # thermal: row 2l (from): p_fr,q_fr ; row 2l+1 (to): p_to,q_to
        th_rows = self.r_thermal + np.repeat(np.arange(2 * nl), 2)
        th_cols = np.stack(
            [self.o_p + ar, self.o_q + ar, self.o_p + nl + ar, self.o_q + nl + ar], 1
        ).ravel()
        rows.append(th_rows)
        cols.append(th_cols)
        self.j_o_thermal = self.j_n_affine
        self.j_thermal_cols = th_cols
        # balance rows
        arcs_at = [[] for _ in range(nb)]
        for l in range(nl):
            arcs_at[f[l]].append(l)  # from-arc index l
            arcs_at[t[l]].append(nl + l)  # to-arc index nl+l
        gens_at = [[] for _ in range(nb)]
        for k in range(ng):
            gens_at[net.gen_bus[k]].append(k)
        self.has_shunt = (net.gs != 0) | (net.bs != 0)
        b_rows, b_cols, b_const = [], [], []
        sh_pos_p, sh_pos_q, sh_bus = [], [], []
        for i in range(nb):
            for part, (ov, og) in enumerate(((self.o_p, self.o_pg), (self.o_q, self.o_qg))):
                r = self.r_bal + 2 * i + part
                for a in arcs_at[i]:
                    b_rows.append(r)
                    b_cols.append(ov + a)
                    b_const.append(1.0)
                for k in gens_at[i]:
                    b_rows.append(r)
                    b_cols.append(og + k)
                    b_const.append(-1.0)
                if self.has_shunt[i]:
                    (sh_pos_p if part == 0 else sh_pos_q).append(len(b_rows))
                    if part == 0:
                        sh_bus.append(i)
                    b_rows.append(r)
                    b_cols.append(self.o_vm + i)
                    b_const.append(0.0)
        self.j_o_bal = self.j_o_thermal + 4 * nl
        rows.append(np.array(b_rows, dtype=np.int64))
        cols.append(np.array(b_cols, dtype=np.int64))
        self.j_bal_const = np.array(b_const)
        self.j_sh_pos_p = np.array(sh_pos_p, dtype=np.int64)
        self.j_sh_pos_q = np.array(sh_pos_q, dtype=np.int64)
        self.sh_bus = np.array(sh_bus, dtype=np.int64)
        self.arcs_at, self.gens_at = arcs_at, gens_at
        # Ohm rows: 5 entries per row: var, vm_i, vm_j, va_i, va_j
        self.j_o_ohm = self.j_o_bal + len(b_rows)
        o_rows = self.r_ohm + np.repeat(np.arange(4 * nl), 5)
        o_cols = np.stack(
            [self.o_var, self.o_vm + self.o_i, self.o_vm + self.o_j, self.o_va + self.o_i, self.o_va + self.o_j], 2
        ).reshape(-1)
        rows.append(o_rows)
        cols.append(o_cols)
        self.j_row = np.concatenate(rows).astype(np.int64) + 1
        self.j_col = np.concatenate(cols).astype(np.int64) + 1

    def _build_hessian_structure(self):
        nb, ng, nl = self.nb, self.ng, self.nl
        rows, cols = [], []
        # quadratic objective: (pg,pg)
        pg = self.o_pg + np.arange(ng)
        rows.append(pg)
        cols.append(pg)
        # thermal rows: (p,p),(q,q) per row
        rows.append(self.j_thermal_cols)
        cols.append(self.j_thermal_cols)
        self.h_o_thermal = ng
        # shunt terms on balance rows: P row then Q row per shunt bus
        self.h_o_shunt = self.h_o_thermal + 4 * nl
        sh = self.o_vm + self.sh_bus
        rows.append(np.repeat(sh, 2))
        cols.append(np.repeat(sh, 2))
        # Ohm rows: 9 entries per row
        self.h_o_ohm = self.h_o_shunt + 2 * sh.shape[0]
        vi, vj = self.o_vm + self.o_i, self.o_vm + self.o_j
        ti, tj = self.o_va + self.o_i, self.o_va + self.o_j
        a = np.stack([vi, vi, vi, vi, vj, vj, ti, ti, tj], 2)
        b = np.stack([vi, vj, ti, tj, ti, tj, ti, tj, tj], 2)
        rows.append(np.maximum(a, b).reshape(-1))
        cols.append(np.minimum(a, b).reshape(-1))
        self.h_row = np.concatenate(rows).astype(np.int64) + 1
        self.h_col = np.concatenate(cols).astype(np.int64) + 1

    # --------------------------------------------------------------- callbacks
    def eval_f(self, x):
        net = self.net
        pg = x[..., self.o_pg : self.o_pg + self.ng]
        return np.sum(net.cost2 * pg * pg + net.cost1 * pg + net.cost0, axis=-1)

    def eval_grad_f(self, x, grad):
        net = self.net
        grad[...] = 0.0
        pg = x[..., self.o_pg : self.o_pg + self.ng]
        grad[..., self.o_pg : self.o_pg + self.ng] = 2.0 * net.cost2 * pg + net.cost1

    def _branch_terms(self, x):
        vi = x[..., self.o_vm + self.o_i]  # [..., nl, 4]
        vj = x[..., self.o_vm + self.o_j]
        th = x[..., self.o_va + self.o_i] - x[..., self.o_va + self.o_j]
        cs, sn = np.cos(th), np.sin(th)
        C = self.oc * cs + self.os * sn
        S = -self.oc * sn + self.os * cs
        return vi, vj, C, S

    def eval_g(self, x, g):
        net, nb, nl = self.net, self.nb, self.nl
        va = x[..., self.o_va : self.o_va + nb]
        vm = x[..., self.o_vm : self.o_vm + nb]
        dth = va[..., net.f_bus] - va[..., net.t_bus]
        g[..., self.r_angU : self.r_angU + nl] = dth
        g[..., self.r_angL : self.r_angL + nl] = dth
        g[..., self.r_ref] = va[..., net.ref_bus]
        p = x[..., self.o_p : self.o_p + 2 * nl]
        q = x[..., self.o_q : self.o_q + 2 * nl]
        s2 = p * p + q * q
        g[..., self.r_thermal : self.r_thermal + 2 * nl : 2] = s2[..., :nl]
        g[..., self.r_thermal + 1 : self.r_thermal + 2 * nl : 2] = s2[..., nl:]
        # balance: sum(p arcs) - sum(pg) + gs vm^2 ; sum(q arcs) - sum(qg) - bs vm^2
        pg = x[..., self.o_pg : self.o_pg + self.ng]
        qg = x[..., self.o_qg : self.o_qg + self.ng]
        bp = _incidence_apply(self.M_arc, p) - _incidence_apply(self.M_gen, pg)
        bq = _incidence_apply(self.M_arc, q) - _incidence_apply(self.M_gen, qg)
        bp += net.gs * vm * vm
        bq -= net.bs * vm * vm
        g[..., self.r_bal : self.r_bal + 2 * nb : 2] = bp
        g[..., self.r_bal + 1 : self.r_bal + 2 * nb : 2] = bq
        vi, vj, C, _ = self._branch_terms(x)
        T = self.oa * vi * vi + vi * vj * C
        g[..., self.r_ohm : self.r_ohm + 4 * nl] = (x[..., self.o_var] - T).reshape(x.shape[:-1] + (4 * nl,))

    def eval_jac_g(self, x, values):
        net, nl = self.net, self.nl
        values[..., : self.j_n_affine] = self.j_affine_vals
        values[..., self.j_o_thermal : self.j_o_thermal + 4 * nl] = 2.0 * x[..., self.j_thermal_cols]
        nbal = self.j_bal_const.shape[0]
        values[..., self.j_o_bal : self.j_o_bal + nbal] = self.j_bal_const
        if self.sh_bus.shape[0]:
            vm_sh = x[..., self.o_vm + self.sh_bus]
            values[..., self.j_o_bal + self.j_sh_pos_p] = 2.0 * net.gs[self.sh_bus] * vm_sh
            values[..., self.j_o_bal + self.j_sh_pos_q] = -2.0 * net.bs[self.sh_bus] * vm_sh
        vi, vj, C, S = self._branch_terms(x)
        blk = np.empty(x.shape[:-1] + (nl, 4, 5))
        blk[..., 0] = 1.0
        blk[..., 1] = -(2.0 * self.oa * vi + vj * C)
        blk[..., 2] = -(vi * C)
        dth = vi * vj * S
        blk[..., 3] = -dth
        blk[..., 4] = dth
        values[..., self.j_o_ohm : self.j_o_ohm + 20 * nl] = blk.reshape(x.shape[:-1] + (20 * nl,))

    def eval_h(self, x, obj_factor, lam, values):
        net, nl, ng = self.net, self.nl, self.ng
        values[..., :ng] = 2.0 * net.cost2 * obj_factor
        lam_th = lam[..., self.r_thermal : self.r_thermal + 2 * nl]
        values[..., self.h_o_thermal : self.h_o_thermal + 4 * nl] = 2.0 * np.repeat(lam_th, 2, axis=-1)
        nsh = self.sh_bus.shape[0]
        if nsh:
            lp = lam[..., self.r_bal + 2 * self.sh_bus]
            lq = lam[..., self.r_bal + 2 * self.sh_bus + 1]
            sh = np.stack([2.0 * net.gs[self.sh_bus] * lp, -2.0 * net.bs[self.sh_bus] * lq], -1)
            values[..., self.h_o_shunt : self.h_o_shunt + 2 * nsh] = sh.reshape(x.shape[:-1] + (2 * nsh,))
        vi, vj, C, S = self._branch_terms(x)
        lo = -lam[..., self.r_ohm : self.r_ohm + 4 * nl].reshape(x.shape[:-1] + (nl, 4))
        blk = np.empty(x.shape[:-1] + (nl, 4, 9))
        blk[..., 0] = 2.0 * self.oa  # vi vi
        blk[..., 1] = C  # vi vj
        blk[..., 2] = vj * S  # vi ti
        blk[..., 3] = -vj * S  # vi tj
        blk[..., 4] = vi * S  # vj ti
        blk[..., 5] = -vi * S  # vj tj
        vvC = vi * vj * C
        blk[..., 6] = -vvC  # ti ti
        blk[..., 7] = vvC  # ti tj
        blk[..., 8] = -vvC  # tj tj
        blk *= lo[..., None]
        values[..., self.h_o_ohm : self.h_o_ohm + 36 * nl] = blk.reshape(x.shape[:-1] + (36 * nl,))


def _incidence_apply(M, v):
    """(M @ v) along the last axis for ``v`` with optional leading batch dims."""
    if v.ndim == 1:
        return M @ v
    flat = v.reshape(-1, v.shape[-1])
    return np.ascontiguousarray((M @ flat.T).T).reshape(v.shape[:-1] + (M.shape[0],))
