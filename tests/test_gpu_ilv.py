"""GPU parity of the interleaved batch layout (csrc/ilv.cuh: G instances per CTA, values batch-innermost -- SURVEY 2.2
K9) against the one-CTA-per-instance team (csrc/ipm.cuh) and the CPU oracle: same algorithm, same inputs, so the
classification must be identical and every returned point must be a KKT point of its own QP to 1e-6 (north_star's bar);
on convex subproblems the steps agree to 1e-6.  Batch sizes that are NOT multiples of G exercise the padding lanes,
`active` masks exercise the per-instance predicates."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "support"))

import closed_loop as cl  # noqa: E402
from oracle.coo import CooMatrix, SymCooMatrix  # noqa: E402
from oracle.subproblem import trust_region_box  # noqa: E402
from sqpsolver_jl_b200 import capi  # noqa: E402
from sqpsolver_jl_b200.host.sqp_trust_region import BatchSqpTR, Parameters  # noqa: E402
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar  # noqa: E402
from sqpsolver_jl_b200.nlp.networks import case9, synth_net  # noqa: E402

pytestmark = pytest.mark.gpu
OK = cl.OK
LAYOUTS = [dict(G=2, threads=512, ctas_per_sm=2), dict(G=4, threads=512, ctas_per_sm=1), dict(G=8, threads=512),
           dict(G=4, threads=512, ctas_per_sm=2)]


def _inputs(net, B, seed, lam_scale):
    pd, qd = net.perturbed_loads(B)
    nlp = AcopfPolar(net, pd=pd, qd=qd)
    rng = np.random.default_rng(seed)
    x = np.clip(np.broadcast_to(nlp.x0, (B, nlp.n)) + 0.02 * rng.standard_normal((B, nlp.n)), nlp.x_L, nlp.x_U)
    lam = lam_scale * rng.standard_normal((B, nlp.m))
    df = np.empty((B, nlp.n)); nlp.eval_grad_f(x, df)
    E = np.empty((B, nlp.m)); nlp.eval_g(x, E)
    dE = np.empty((B, nlp.nnz_jac_coo)); nlp.eval_jac_g(x, dE)
    hv = np.empty((B, nlp.nnz_hess_coo)); nlp.eval_h(x, 1.0, lam, hv)
    return nlp, x, dE, hv, df, E


def _solve_all(nlp, B, x, dE, hv, df, E, delta, layout, active=None):
    eng = capi.Engine(0)
    try:
        eng.set_layout(**layout)
        eng.setup_nlp(nlp.n, nlp.m, nlp.num_linear_constraints, nlp.j_row, nlp.j_col, nlp.h_row, nlp.h_col, nlp.x_L, nlp.x_U,
                      nlp.g_L, nlp.g_U, batch=B)
        eng.update_nlp(dE, hv, df, E)
        out = {}
        for name, ph, d in (("qp", capi.PHASE_QP, delta), ("fr", capi.PHASE_FR, delta), ("lp", capi.PHASE_LP, np.inf)):
            r = eng.solve_tr(ph, x, d, active=active)
            out[name] = [np.array(v, copy=True) for v in r]
        return out
    finally:
        eng.close()


@pytest.mark.parametrize("layout", LAYOUTS, ids=lambda l: "G%d_%d_%d" % (l["G"], l["threads"], l.get("ctas_per_sm", 0)))
@pytest.mark.parametrize("case", ["case9", "case118"])
def test_interleaved_layout_matches_cta_team_and_kkt(built_lib, layout, case):
    net, B, delta, lam_scale = (case9(), 37, 0.3, 5.0) if case == "case9" else (synth_net(118, 186, 54, seed=118), 70, 2.0, 20.0)
    nlp, x, dE, hv, df, E = _inputs(net, B, 7, lam_scale)
    ref = _solve_all(nlp, B, x, dE, hv, df, E, delta, dict(G=1))
    got = _solve_all(nlp, B, x, dE, hv, df, E, delta, layout)
    J = CooMatrix(nlp.j_row, nlp.j_col, nlp.m, nlp.n)
    H = SymCooMatrix(nlp.h_row, nlp.h_col, nlp.n)
    gL = np.broadcast_to(nlp.g_L, (B, nlp.m)); gU = np.broadcast_to(nlp.g_U, (B, nlp.m))
    for ph in ("qp", "fr", "lp"):
        st_r, st_g = ref[ph][5], got[ph][5]
        cls = lambda s: np.where(np.isin(s, OK), 1, np.where(np.isin(s, cl.INFEAS), 2, 3))
        assert np.array_equal(cls(st_r), cls(st_g)), (ph, st_r, st_g)
        assert (cls(st_g) != 3).all(), (ph, st_g)
        it_r, it_g = ref[ph][6]["ipm_iters"], got[ph][6]["ipm_iters"]
        assert np.abs(it_r.astype(int) - it_g.astype(int)).max() <= 6, (ph, it_r, it_g)  # same algorithm, other summation order
        okm = np.isin(st_g, OK)
        if ph == "qp":
            worst = 0.0
            for b in np.nonzero(okm)[0]:
                J.fill(dE[b]); H.fill(hv[b])
                lb, ub = trust_region_box(nlp.x_L - x[b], nlp.x_U - x[b], delta)
                worst = max(worst, cl.scaled_kkt(H.to_scipy(), df[b], J.to_scipy(), gL[b] - E[b], gU[b] - E[b], lb, ub, got[ph][0][b],
                                                 got[ph][1][b], got[ph][2][b] + got[ph][3][b]))
            assert worst <= 1e-6, worst
        if ph == "lp":  # strictly convex projection: unique solution
            assert np.abs(got[ph][0][okm] - ref[ph][0][okm]).max() <= 1e-6
        if ph == "fr":  # LP: same optimal sum of slacks
            a, b_ = ref[ph][4][okm].sum(axis=1), got[ph][4][okm].sum(axis=1)
            assert np.abs(a - b_).max() <= 1e-6 * max(1.0, np.abs(a).max())
        # infeasible instances are zero-filled (collect_solution! :551-555)
        inf = np.isin(st_g, cl.INFEAS)
        assert not got[ph][0][inf].any() and not got[ph][1][inf].any()


def test_interleaved_layout_active_mask(built_lib):
    """Instances with active[b] == 0 are skipped and their outputs left untouched (sqpqp.h), also inside a group."""
    net, B = case9(), 11
    nlp, x, dE, hv, df, E = _inputs(net, B, 3, 0.0)
    layout = dict(G=4, threads=512, ctas_per_sm=1)
    full = _solve_all(nlp, B, x, dE, hv, df, E, 0.3, layout)
    act = np.array([1, 0, 0, 1, 0, 0, 0, 0, 1, 1, 0], np.int32)   # group 1 entirely inactive
    part = _solve_all(nlp, B, x, dE, hv, df, E, 0.3, layout, active=act)
    on = act.astype(bool)
    assert np.array_equal(part["qp"][0][on], full["qp"][0][on])            # bit-identical: deterministic reductions
    assert np.array_equal(part["qp"][1][on], full["qp"][1][on])
    assert not part["qp"][0][~on].any() and (part["qp"][5][~on] == 0).all()  # untouched (freshly zeroed buffers)


def test_interleaved_layout_is_bit_reproducible(built_lib):
    net, B = case9(), 9
    nlp, x, dE, hv, df, E = _inputs(net, B, 5, 5.0)
    a = _solve_all(nlp, B, x, dE, hv, df, E, 0.3, dict(G=4, threads=512))
    b = _solve_all(nlp, B, x, dE, hv, df, E, 0.3, dict(G=4, threads=512))
    assert all(np.array_equal(u, v) for u, v in zip(a["qp"][:5], b["qp"][:5]))


def test_batched_sqp_with_interleaved_layout_matches_individual_oracle_runs(built_lib):
    """BASELINE configs[4] in miniature through the interleaved path: 6 perturbed-load case9 instances (not a multiple
    of G) against six oracle SQP runs."""
    from oracle.sqp_tr import Parameters as OParams, SqpTROracle
    net = case9()
    B = 6
    pd, qd = net.perturbed_loads(B, rel_sigma=0.05, seed=1234)
    kw = dict(max_iter=100, init_mu=1e4)
    bt = BatchSqpTR(AcopfPolar(net, pd=pd, qd=qd), B, Parameters(**kw), layout=dict(G=4, threads=512)).run()
    for b in range(B):
        o = SqpTROracle(AcopfPolar(net, pd=pd[b], qd=qd[b]), OParams(**kw)).run()
        assert bt.status[b] == o.status == 0
        assert abs(bt.obj_val[b] - o.obj_val) <= 1e-6 * abs(o.obj_val)
    bt.close()
