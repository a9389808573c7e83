"""GPU parity of the resident CTA team that streams its index programs through the shared-memory ring (csrc/chol.cuh:
ring_run_segment, cp.async.bulk + mbarrier) against the slot-list code it replaces and against the oracle's QPs.

Reference: the solve at subproblem_JuMP.jl:178 (JuMP.optimize! of the QP subproblem); QP data of sqp_trust_region.jl:314-331.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "support"))

import closed_loop as cl  # noqa: E402
from sqpsolver_jl_b200 import capi  # noqa: E402
from sqpsolver_jl_b200.host.sqp_trust_region import BatchSqpTR, Parameters  # noqa: E402
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar  # noqa: E402
from sqpsolver_jl_b200.nlp.networks import case9, synth_net  # noqa: E402
from sqpsolver_jl_b200.nlp.toy import ToyExample  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(HERE, "golden")
OK = cl.OK


def _setup(eng, nlp, batch=1):
    eng.setup_nlp(nlp.n, nlp.m, nlp.num_linear_constraints, nlp.j_row, nlp.j_col, nlp.h_row, nlp.h_col,
                  nlp.x_L, nlp.x_U, nlp.g_L, nlp.g_U, batch=batch)


@pytest.mark.parametrize("name,make", [("toy_example", ToyExample), ("case9_default", lambda: AcopfPolar(case9())),
                                       ("case118_first", lambda: AcopfPolar(synth_net(118, 186, 54, 118)))])
def test_ring_matches_slot_lists_on_golden_subproblems(built_lib, name, make):
    """Same subproblem, ring on / ring off: same classification, both KKT points to 1e-6, same optimal value, and nearly the
    same number of interior-point iterations (the two programs sum the same products in a different order)."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    nlp = make()
    res = {}
    for ring in (0, 1):
        eng = capi.Engine(0)
        eng.set_layout(ring=ring)
        _setup(eng, nlp)
        eng.set_options(warm_start=0)
        out = []
        for k in range(g["qp_status"].shape[0]):
            eng.update_nlp(g["qp_dE"][k], g["qp_h_val"][k], g["qp_df"][k], g["qp_E"][k])
            fr = bool(g["qp_fr"][k])
            p, lam, mxL, mxU, slack, st, info = eng.solve_tr(capi.PHASE_FR if fr else capi.PHASE_QP, g["qp_x"][k], g["qp_Delta"][k])
            out.append((p[0].copy(), lam[0].copy(), (mxL[0] + mxU[0]).copy(), int(st[0]), int(info[0]["ipm_iters"]), eng.last_solve_kernel))
        res[ring] = out
        eng.close()
    assert any(o[5].endswith("+ring") for o in res[0]), [o[5] for o in res[0]][:3]
    assert not any(o[5].endswith("+ring") for o in res[1])
    for k, (a, b) in enumerate(zip(res[0], res[1])):
        assert (a[3] in OK) == (b[3] in OK) and (a[3] in cl.INFEAS) == (b[3] in cl.INFEAS), (name, k, a[3], b[3])
        if a[3] not in OK or bool(g["qp_fr"][k]):
            continue
        t = {"dE": g["qp_dE"][k], "h_val": g["qp_h_val"][k], "df": g["qp_df"][k], "E": g["qp_E"][k], "x": g["qp_x"][k], "Delta": g["qp_Delta"][k]}
        P, q, A, rl, ru, xl, xu = cl.qp_of_trace(nlp, t)
        for o in (a, b):
            assert cl.scaled_kkt(P, q, A, rl, ru, xl, xu, o[0], o[1], o[2]) <= 1e-6, (name, k)
        assert abs(a[4] - b[4]) <= max(3, a[4] // 4), (name, k, a[4], b[4])
        oa = 0.5 * a[0] @ (P @ a[0]) + q @ a[0]
        ob = 0.5 * b[0] @ (P @ b[0]) + q @ b[0]
        if np.abs(a[0] - b[0]).max() <= 1e-6 * max(1.0, np.abs(b[0]).max()):
            assert abs(oa - ob) <= 1e-6 * max(1.0, abs(ob))


def test_ring_is_bit_reproducible_and_batched(built_lib):
    """148 perturbed case118-shaped instances (one CTA per SM, ring on): two runs give identical bits; every instance is a
    KKT point of its own QP; instance 0 equals the single-instance solve bit for bit (no cross-instance state)."""
    B = 148
    net = synth_net(118, 186, 54, seed=118)
    pd, qd = net.perturbed_loads(B)
    nlp = AcopfPolar(net, pd=pd, qd=qd)
    runs = []
    for _ in range(2):
        bt = BatchSqpTR(nlp, B, Parameters(max_iter=3, init_mu=1e5))
        bt.trace = []
        bt.trace_instances = {0, 77, 147}
        bt.run()
        assert bt.optimizer.engine.last_solve_kernel.endswith("+ring"), bt.optimizer.engine.last_solve_kernel
        runs.append((bt.x.copy(), bt.lam.copy(), [t for t in bt.trace]))
        bt.close()
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])
    for b in (0, 77, 147):
        tr = [dict(t) for t in runs[0][2] if t["b"] == b]
        for t in tr:
            t["b"] = None
        s = cl.check_trace(AcopfPolar(net, pd=pd[b], qd=qd[b]), tr, oracle_every=2)
        assert s["worst_kkt"] <= 1e-6 and s["marginal_mismatch"] <= 1, (b, s)


def test_handoff_to_the_resident_launch_matches_the_uninterrupted_solve(built_lib):
    """Hand-off (launch_solve): the throughput launch stops every instance after its iteration quota, saves the loop state, and
    the resident ring launch continues it.  With a quota of 10 nearly every instance changes kernels mid-solve; the result must
    be the solve the uninterrupted launch gives: same classification, KKT points to 1e-6, the same step where the QP has one
    solution, nearly the same iteration count (the two kernels sum the factorisation in a different order)."""
    B = 200  # more instances than SMs: two CTAs per SM, values in L2
    net = synth_net(118, 186, 54, seed=118)
    pd, qd = net.perturbed_loads(B)
    nlp = AcopfPolar(net, pd=pd, qd=qd)
    res = {}
    for quota in (0, 10):
        bt = BatchSqpTR(nlp, B, Parameters(max_iter=2, init_mu=1e5), layout=dict(handoff=quota))
        bt.trace = []
        bt.trace_instances = {0, 57, 199}
        recs = []
        orig = bt.optimizer._solve

        def hook(phase, x_k, delta, E_override=None, active=None, _o=orig, _r=recs, _bt=bt):
            out = _o(phase, x_k, delta, E_override, active)
            info = _bt.optimizer.last_info
            _r.append((phase, out[0].copy(), out[-1].copy(), info["ipm_iters"].copy(), _bt.optimizer.engine.last_solve_kernel))
            return out

        bt.optimizer._solve = hook
        bt.run()
        res[quota] = (recs, [dict(t) for t in bt.trace])
        bt.close()
    assert all("+handoff" not in r[4] for r in res[0][0])
    assert any("+handoff" in r[4] for r in res[10][0]), [r[4] for r in res[10][0]]
    for a, b in zip(res[0][0], res[10][0]):
        assert a[0] == b[0]
        sa, sb = np.asarray(a[2]), np.asarray(b[2])
        assert np.array_equal(np.isin(sa, OK), np.isin(sb, OK)), (sa[~np.isin(sa, OK)], sb[~np.isin(sb, OK)])
        assert np.abs(a[3].astype(int) - b[3].astype(int)).max() <= 4, (a[3][:8], b[3][:8])
        ok = np.isin(sa, OK)
        if a[0] == capi.PHASE_LP:  # strictly convex projection: one solution
            assert np.abs(a[1][ok] - b[1][ok]).max() <= 1e-6 * max(1.0, np.abs(a[1][ok]).max())
    for b in (0, 57, 199):
        tr = [t for t in res[10][1] if t["b"] == b]
        for t in tr:
            t["b"] = None
        s = cl.check_trace(AcopfPolar(net, pd=pd[b], qd=qd[b]), tr, oracle_every=2)
        assert s["worst_kkt"] <= 1e-6 and s["marginal_mismatch"] <= 1, (b, s)


def test_launch_order_and_two_stage_launch_do_not_change_the_results(built_lib):
    """sqpqp_set_launch_order permutes which CTA slot runs which instance; the two-stage launch (iteration quota, loop state
    saved, every unfinished instance resumed by a second launch of the SAME kernel, in index order or in the order the
    device ranks from the saved states) interrupts every solve once.  Neither may change a bit of the results: the same code
    runs every instance on the same data (profiles/r02_tuning.md section 8: an oracle order is worth 15-24 % of a multi-wave
    launch, but nothing known before or early in a solve predicts its length)."""
    B = 320  # more than 2 x 148 resident CTAs: the order matters for who starts first
    net = synth_net(118, 186, 54, seed=118)
    pd, qd = net.perturbed_loads(B)
    nlp = AcopfPolar(net, pd=pd, qd=qd)
    bt = BatchSqpTR(nlp, B, Parameters(max_iter=2, init_mu=1e5))
    eng = bt.optimizer.engine
    recs = {}
    orig = bt.optimizer._solve
    state = {"n": 0}

    def hook(phase, x_k, delta, E_override=None, active=None):
        state["n"] += 1
        if phase != capi.PHASE_QP or state["n"] < 2:
            return orig(phase, x_k, delta, E_override, active)
        rng = np.random.default_rng(5)

        def run(tag):
            out = orig(phase, x_k, delta, E_override, active)
            recs[tag] = (out[0].copy(), out[1].copy(), np.asarray(out[-1]).copy(), bt.optimizer.last_info.copy(), eng.last_solve_kernel)

        run("base")
        eng.set_launch_order(rng.permutation(B))
        run("permuted")
        eng.set_launch_order(None)
        for mode, tag in ((3, "two_stage"), (1, "two_stage_ranked")):
            eng.set_layout(handoff=9, handoff_mode=mode)
            run(tag)
        eng.set_layout(handoff=-1, handoff_mode=0)
        return orig(phase, x_k, delta, E_override, active)

    bt.optimizer._solve = hook
    bt.run()
    bt.close()
    assert set(recs) == {"base", "permuted", "two_stage", "two_stage_ranked"}
    p0, l0, s0, i0, k0 = recs["base"]
    assert i0["ipm_iters"].min() > 9  # every instance is interrupted by the quota of the two-stage runs
    assert "+resume" in recs["two_stage"][4] and "ranked" in recs["two_stage_ranked"][4], (recs["two_stage"][4], recs["two_stage_ranked"][4])
    for tag in ("permuted", "two_stage", "two_stage_ranked"):
        p, lam, st, info, _ = recs[tag]
        assert np.array_equal(st, s0), tag
        assert np.array_equal(info["ipm_iters"], i0["ipm_iters"]) and np.array_equal(info["chol_factorizations"], i0["chol_factorizations"]), tag
        assert np.array_equal(p, p0) and np.array_equal(lam, l0), (tag, np.abs(p - p0).max())
    with pytest.raises(capi.SqpQpError):
        eng2 = capi.Engine()
        try:
            _setup(eng2, AcopfPolar(case9()), batch=4)
            eng2.set_launch_order([0, 1, 1, 3])  # not a permutation
        finally:
            eng2.close()
