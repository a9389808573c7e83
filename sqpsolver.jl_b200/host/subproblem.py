"""Host-side mirror of the reference's sub-optimizer interface, backed by the device engine.

``QpDevice`` plays the role of ``QpJuMP <: AbstractSubOptimizer``
(src/algorithms/subproblem_JuMP.jl:1-24) with the same method set and return tuples:

    create_model!(qp, Delta)            -> create_model(delta)
    sub_optimize!(qp, x_k, Delta)       -> sub_optimize(x_k, delta)        (:127-183)
    sub_optimize_FR!(qp, x_k, Delta)    -> sub_optimize_FR(x_k, delta)     (:352-393)
    sub_optimize_lp(optimizer, A, ...)  -> sub_optimize_lp(x_k)            (:185-244)
    (SOC solve, sqp_trust_region.jl:341-360) -> sub_optimize_soc(x_k, delta, E_soc)

each returning ``(p, lambda, mult_x_U, mult_x_L, p_slack, status)`` exactly as
``collect_solution!`` (:514-563) does -- but for a whole batch of instances that share one
sparsity pattern (arrays carry a leading batch axis; a single NLP is a batch of 1).
Instead of rewriting a JuMP model coefficient by coefficient (:465-512), the raw COO
value arrays of the evaluator go to the GPU through ``update`` (pinned async copies inside
the C-ABI) and the QP is assembled there.

Everything numerical happens in ``csrc/`` through :mod:`..capi`; there is no CPU path.
"""
from __future__ import annotations

import numpy as np

from .. import capi

OK_STATUSES = (capi.MOI_OPTIMAL, 7, capi.MOI_ALMOST_LOCALLY_SOLVED, capi.MOI_LOCALLY_SOLVED)
INFEASIBLE_STATUSES = (capi.MOI_INFEASIBLE, capi.MOI_LOCALLY_INFEASIBLE)


class QpDevice:
    def __init__(self, nlp, batch: int = 1, device: int = 0, engine_options: dict | None = None, layout: dict | None = None):
        self.nlp = nlp
        self.batch = batch
        self.engine = capi.Engine(device)
        if engine_options:
            self.engine.set_options(**engine_options)
        if layout:
            self.engine.set_layout(**layout)
        self.created = False
        self.stats = {"solves": 0, "instance_solves": 0, "admm_iters": 0, "cg_iters": 0, "polish_cg_iters": 0,
                      "checks": 0, "solve_ms": 0.0, "polished": 0}

    def create_model(self, delta=None):
        nlp = self.nlp
        self.engine.setup_nlp(nlp.n, nlp.m, nlp.num_linear_constraints, nlp.j_row, nlp.j_col, nlp.h_row, nlp.h_col,
                              nlp.x_L, nlp.x_U, nlp.g_L, nlp.g_U, batch=self.batch)
        self.created = True

    def update(self, dE, h_val, df, E):
        """QpData refresh (sqp.jl:66-79) + eval_Jacobian!/Hessian scatter (sqp.jl:92-117)."""
        self.engine.update_nlp(dE, h_val, df, E)

    def _solve(self, phase, x_k, delta, E_override=None, active=None):
        if phase == capi.PHASE_MIXED:  # active = (mask of the QP phase, mask of the restoration phase)
            p, lam, mxL, mxU, slack, status, info = self.engine.solve_tr_mixed(x_k, delta, active[0], active[1])
            active = np.asarray(active[0], bool) | np.asarray(active[1], bool)
        else:
            p, lam, mxL, mxU, slack, status, info = self.engine.solve_tr(phase, x_k, delta, E_override, active)
        sel = slice(None) if active is None else np.asarray(active, bool)
        st = self.stats
        st["solves"] += 1
        st["instance_solves"] += int(self.batch if active is None else np.sum(sel))
        for k_, f_ in (("admm_iters", "admm_iters"), ("cg_iters", "cg_iters"), ("polish_cg_iters", "polish_cg_iters"),
                       ("checks", "checks"), ("polished", "polished")):
            st[k_] += int(info[f_][sel].sum())
        st["solve_ms"] += self.engine.last_solve_ms
        self.last_info = info
        return p, lam, mxU, mxL, slack, status

    def sub_optimize(self, x_k, delta, active=None):
        return self._solve(capi.PHASE_QP, x_k, delta, None, active)

    def sub_optimize_FR(self, x_k, delta, active=None):
        return self._solve(capi.PHASE_FR, x_k, delta, None, active)

    def sub_optimize_mixed(self, x_k, delta, active_qp, active_fr):
        """One round of compute_step! (sqp_trust_region.jl:370-380) over a batch with instances in BOTH phases: sub_optimize!
        for `active_qp`, sub_optimize_FR! for `active_fr` (disjoint masks), one call -- the two launches run side by side."""
        return self._solve(capi.PHASE_MIXED, x_k, delta, None, (active_qp, active_fr))

    def sub_optimize_soc(self, x_k, delta, E_soc, active=None):
        return self._solve(capi.PHASE_SOC, x_k, delta, E_soc, active)

    def sub_optimize_lp(self, x_k, active=None):
        p, lam, mxU, mxL, _, status = self._solve(capi.PHASE_LP, x_k, np.full(self.batch, np.inf), None, active)
        return p, lam, mxU, mxL, status

    def close(self):
        self.engine.close()
