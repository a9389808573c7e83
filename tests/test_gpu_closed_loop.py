"""GPU parity, whole trajectories: closed-loop replay of the device SQP runs against the CPU oracle on every BASELINE
config (tests/support/closed_loop.py), the `lb > ub` box fallback on the device, and boundary B1 validated with the
JuMP model the reference really builds (tests/support/jump_replay.py).

Reference: sqp_trust_region.jl:98-223 (run!), subproblem_JuMP.jl:36-183, 352-393, 432-563, test/runtests.jl:12-14.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "support"))

import closed_loop as cl  # noqa: E402
from jump_replay import JumpReplay  # noqa: E402
from oracle import qp_solver as qs  # noqa: E402
from oracle.sqp_tr import Parameters as OParams, SqpTROracle  # noqa: E402
from oracle.subproblem import trust_region_box  # noqa: E402
from sqpsolver_jl_b200 import capi  # noqa: E402
from sqpsolver_jl_b200.host.sqp_trust_region import BatchSqpTR, Parameters, SqpTR  # noqa: E402
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar  # noqa: E402
from sqpsolver_jl_b200.nlp.networks import case9, synth_net  # noqa: E402
from sqpsolver_jl_b200.nlp.toy import ReadmeToy, ToyExample  # noqa: E402

pytestmark = pytest.mark.gpu
OK = cl.OK


# ------------------------------------------------------------------------------------------------ closed loop
@pytest.mark.parametrize("name,make,kw,every", [
    ("toy", ToyExample, dict(max_iter=100), 1),
    ("readme_toy", ReadmeToy, dict(max_iter=100), 1),
    ("case9_mu1e4", lambda: AcopfPolar(case9()), dict(max_iter=100, init_mu=1e4), 1),
    # the reference's own defaults (parameters.jl:17-29: init_mu = 1) as test/opf.jl:18-23 runs them
    ("case9_default", lambda: AcopfPolar(case9()), dict(max_iter=100), 1),
    ("case9_soc", lambda: AcopfPolar(case9()), dict(max_iter=100, init_mu=1e4, use_soc=True), 1),
    # BASELINE configs[2]: case118-shaped network, trust-region variant, 100 SQP iterations, every subproblem checked
    ("case118", lambda: AcopfPolar(synth_net(118, 186, 54, 118)), dict(max_iter=100, init_mu=1e5), 1),
    ("case118_default", lambda: AcopfPolar(synth_net(118, 186, 54, 118)), dict(max_iter=40), 1),
])
def test_closed_loop_single_instance(built_lib, name, make, kw, every):
    nlp = make()
    dev = SqpTR(make(), Parameters(**kw))
    trace = []
    dev.run(trace=trace)
    dev.close()
    assert len(trace) >= 1
    s = cl.check_trace(nlp, trace, oracle_every=every)
    print(name, "status", dev.status, "iters", dev.iter, s)
    assert s["worst_kkt"] <= 1e-6
    assert s["marginal_mismatch"] <= max(2, s["n"] // 10), s   # marginal (violation < 1e-6) subproblems only, and few
    assert s["oracle_solved"] >= 1
    if name in ("toy", "readme_toy", "case9_mu1e4", "case9_soc"):
        assert dev.status == 0


def test_closed_loop_batch_case118(built_lib):
    """BASELINE configs[4]: 1024 perturbed-load case118-shaped instances in one batch; 8 sampled instances are replayed
    subproblem by subproblem against the oracle over the first 12 SQP rounds."""
    B = 1024
    net = synth_net(118, 186, 54, seed=118)
    pd, qd = net.perturbed_loads(B)
    nlp = AcopfPolar(net, pd=pd, qd=qd)
    bt = BatchSqpTR(nlp, B, Parameters(max_iter=12, init_mu=1e5))
    sample = [0, 17, 128, 341, 512, 682, 900, 1023]
    bt.trace = []
    bt.trace_instances = set(sample)
    bt.run()
    bt.close()
    info_st = np.unique(bt.status, return_counts=True)
    print("batch status", info_st)
    for b in sample:
        tr = [t for t in bt.trace if t["b"] == b]
        assert len(tr) >= 10
        one = AcopfPolar(net, pd=pd[b], qd=qd[b])
        for t in tr:
            t["b"] = None  # bounds of the single-instance NLP are 1-D
        s = cl.check_trace(one, tr, oracle_every=1)
        print("instance", b, s)
        assert s["worst_kkt"] <= 1e-6 and s["marginal_mismatch"] <= 2


def test_closed_loop_case2000(built_lib):
    """BASELINE configs[3]: the ~2000-bus network, 12 SQP iterations; status / storage convention / box on every subproblem,
    HiGHS verdicts on the first two infeasible QPs, the oracle on the first two restoration LPs (a 2000-bus LP takes the
    CPU oracle ~20 s, so the test stays within a couple of minutes)."""
    make = lambda: AcopfPolar(synth_net(2000, 3000, 400, 2000))
    nlp = make()
    dev = SqpTR(make(), Parameters(max_iter=12, init_mu=1e5))
    trace = []
    dev.run(trace=trace)
    dev.close()
    assert len(trace) >= 12
    # the first iterations alternate between an infeasible QP and its restoration LP
    fr_ids = [k for k, t in enumerate(trace) if t["fr"] and int(t["status"]) in OK][:2]
    s = cl.check_trace(nlp, trace, oracle_on=set(fr_ids), feas_checks=2)
    print("case2000", dev.status, s)
    assert s["worst_kkt"] <= 1e-6 and s["marginal_mismatch"] <= 1 and s["oracle_solved"] >= 2


# ------------------------------------------------------------------------------------------------ box fallback
def test_trust_region_box_fallback_on_device(engine):
    """set_trust_region! with x_k OUTSIDE its bounds (subproblem_JuMP.jl:441-444): lb > ub -> lb = max(-D, min(0, v_lb)),
    ub = min(D, max(0, v_ub)).  A separable strictly convex QP makes the box the only thing that decides p."""
    n, m = 6, 1
    idx = np.arange(1, n + 1)
    x_L, x_U = np.full(n, -1.0), np.full(n, 1.0)
    engine.setup_nlp(n, m, 1, np.ones(n, np.int64), idx, idx, idx, x_L, x_U, np.array([-1e3]), np.array([1e3]))
    x_k = np.array([3.0, -4.0, 0.5, 1.0, -1.0, 10.0])   # 0, 1, 5 violate their bounds
    delta = 0.5
    df = np.array([-100.0, 100.0, -100.0, 100.0, -100.0, 100.0])
    engine.update_nlp(np.ones(n), np.ones(n), df, np.zeros(m))
    p, lam, mxL, mxU, _, st, info = engine.solve_tr(capi.PHASE_QP, x_k, delta)
    assert st[0] in OK
    lb, ub = trust_region_box(x_L - x_k, x_U - x_k, delta)
    # x0: v = [-4,-2] -> lb=-0.5 > ub=-2 -> fallback [-0.5, 0];  x1: v = [3,5] -> fallback [0, 0.5];  x5: v=[-11,-9] -> [-0.5,0]
    assert np.allclose(lb, [-0.5, 0.0, -0.5, -0.5, 0.0, -0.5]) and np.allclose(ub, [0.0, 0.5, 0.5, 0.0, 0.5, 0.0])
    want = np.clip(-df, lb, ub)  # min 1/2 p'p + df'p over the box (row inactive)
    assert np.abs(p[0] - want).max() <= 1e-7, (p[0], want)
    assert (p[0] >= lb - 1e-9).all() and (p[0] <= ub + 1e-9).all()
    res = qs.solve_qp(np.eye(n), df, np.ones((1, n)), np.array([-1e3]), np.array([1e3]), lb, ub)
    assert np.abs(p[0] - res.x).max() <= 1e-6
    assert np.abs((mxL[0] + mxU[0]) - res.col_dual).max() <= 1e-6 * max(1.0, np.abs(res.col_dual).max())


# ------------------------------------------------------------------------------------------------ boundary B1
@pytest.mark.parametrize("make,kw", [(ToyExample, dict(max_iter=100)), (ReadmeToy, dict(max_iter=100)),
                                     (lambda: AcopfPolar(case9()), dict(max_iter=100, init_mu=1e4))])
def test_generic_lane_with_the_jump_model_of_the_reference(built_lib, make, kw):
    """A whole SQP solve through sqpqp_qp_setup / sqpqp_qp_solve with the model `create_model!` / `sub_optimize!` /
    `sub_optimize_FR!` / `modify_constraints!` build (n + S columns, slacks fixed / freed through column bounds, paired
    range rows at m + k, objective rebuilt per solve), flattened the way SqpQpB200.copy_to does and read back the way
    collect_solution! does -- against the NLP lane (QpDevice) on the same problem."""
    eng = capi.Engine(0)
    try:
        replay = []

        def factory(data):
            r = JumpReplay(data, eng)
            replay.append(r)
            return r

        log_g, log_n = [], []
        gen = SqpTROracle(make(), OParams(**kw), sub_factory=factory).run(log_g)
        nlp_lane = SqpTR(make(), Parameters(**kw)).run(log_n)
        nlp_lane.close()
        assert gen.status == nlp_lane.status == 0
        assert abs(gen.obj_val - nlp_lane.obj_val) <= 1e-6 * max(1.0, abs(nlp_lane.obj_val))
        assert np.abs(gen.x - nlp_lane.x).max() <= 1e-5 * max(1.0, np.abs(nlp_lane.x).max())
        # same trajectory, iteration by iteration (objective, trust region, step length, phase) for as long as the two
        # lanes see the same QP: the two formulations (n + S columns with fixed slacks and paired rows vs n columns and
        # two-sided rows) take different interior-point paths, and on a NONCONVEX subproblem (case9, indefinite Lagrangian
        # Hessian) they may stop at different local solutions -- from there on the runs differ and only meet again at
        # the optimum.  The toy problems must agree over the whole run.
        same = 0
        for a, b in zip(log_g, log_n):
            if not (a["fr"] == b["fr"] and a["accept"] == b["accept"] and
                    all(abs(a[key] - b[key]) <= 1e-5 * max(1.0, abs(b[key])) for key in ("f", "Delta", "pinf", "inf_pr"))):
                break
            same += 1
        print("generic vs NLP lane: identical iterations", same, "of", len(log_g), len(log_n))
        if len(log_n) <= 6:
            assert same == len(log_g) == len(log_n)
        # case9: the multipliers of its QPs are not unique (LICQ fails), the two formulations return different points of
        # the multiplier face, the Hessian of the next iteration is evaluated with them (sqp.jl:93) and the runs separate
        # from iteration 2 on -- both reach the same optimum (asserted above)
        r = replay[0]
        # every QP / restoration subproblem went through the generic lane (n_qp also counts the start-point projection, which
        # the oracle driver solves on the CPU when the start violates a linear row or a bound: subproblem_JuMP.jl:185-244)
        assert len(replay) == 1 and r.n_solves in (gen.n_qp, gen.n_qp - 1)
        # the device structure is rebuilt only when the pattern of the model changes (QP <-> restoration LP objective)
        assert r.n_setups <= 2 * sum(1 for a, b in zip(log_g, log_g[1:]) if a["fr"] != b["fr"]) + 2
    finally:
        eng.close()


def test_jump_model_duals_land_on_the_right_rows(engine):
    """One QP with every row kind (==, >=, <=, two-sided; linear and nonlinear) through the replayed JuMP model: the
    multipliers mapped back by collect_solution! (:531-550) must equal the oracle's on the reference's row numbering."""
    from oracle.subproblem import QpData, QpOracle
    rng = np.random.default_rng(4)
    n, m, ml = 7, 8, 3
    A = rng.standard_normal((m, n))
    xs = rng.uniform(-0.3, 0.3, n)
    E = np.zeros(m)
    Ax = A @ xs
    c_lb = np.array([Ax[0], Ax[1] - 0.1, -np.inf, Ax[3], Ax[4] - 0.2, -np.inf, Ax[6] + 0.0, Ax[7] - 0.05])
    c_ub = np.array([Ax[0], np.inf, Ax[2] + 0.1, Ax[3], np.inf, Ax[5] + 0.02, Ax[6] + 0.3, Ax[7] + 0.05])
    M = rng.standard_normal((n, n))
    Q = M @ M.T + np.eye(n)
    c = 3.0 * rng.standard_normal(n)
    import scipy.sparse as sp
    data = QpData(sp.csr_matrix(Q), c, sp.csr_matrix(A), E, c_lb, c_ub, np.full(n, -0.5), np.full(n, 0.5), ml)
    ora = QpOracle(data); ora.create_model(1.0)
    xo, lo, uo, Lo, _, so = ora.sub_optimize(np.zeros(n), 1.0)
    rep = JumpReplay(data, engine); rep.create_model(1.0)
    xg, lg, ug, Lg, ps, sg = rep.sub_optimize(np.zeros(n), 1.0)
    assert so in qs.OK_STATUSES and sg in qs.OK_STATUSES
    assert rep.ncol == n + (m - ml) + 3          # u per nonlinear row, v for the == row and the two two-sided rows
    assert len(rep.constr) == m + 2              # the `<=` halves of the two-sided rows appended at m + k
    assert set(ps.keys()) == set(range(ml + 1, m + 1)) and all(v == 0.0 for vs in ps.values() for v in vs)
    assert np.abs(xg - xo).max() <= 1e-6
    assert np.abs(lg - lo).max() <= 1e-6 * max(1.0, np.abs(lo).max())
    assert np.abs((ug + Lg) - (uo + Lo)).max() <= 1e-6 * max(1.0, np.abs(uo + Lo).max())


# ------------------------------------------------------------------------------------------------ driver parity
def _prefix_equal(la, lb, keys, rtol):
    same = 0
    for a, b in zip(la, lb):
        if a["fr"] != b["fr"] or any(abs(a[k] - b[k]) > rtol * max(1.0, abs(b[k])) for k in keys):
            break
        same += 1
    return same


@pytest.mark.parametrize("name,make,kw,whole", [
    ("toy", ToyExample, dict(max_iter=100), True),
    ("readme_toy", ReadmeToy, dict(max_iter=100), True),
    ("case9_mu1e4", lambda: AcopfPolar(case9()), dict(max_iter=100, init_mu=1e4), True),
    # reference defaults (init_mu = 1): the run two different QP solvers cannot be compared on (degenerate restoration LPs)
    ("case9_default", lambda: AcopfPolar(case9()), dict(max_iter=100), True),
    ("case9_soc", lambda: AcopfPolar(case9()), dict(max_iter=100, init_mu=1e4, use_soc=True), True),
    # BASELINE configs[2]: does not converge (reference algorithm as coded); the runs must agree while rounding noise
    # (1e-12 relative, different summation order of the norms) has not been amplified by the nonconvex subproblems
    ("case118", lambda: AcopfPolar(synth_net(118, 186, 54, 118)), dict(max_iter=30, init_mu=1e5), False),
])
def test_trust_region_driver_parity_with_shared_subsolver(built_lib, name, make, kw, whole):
    """SQP-TR: the oracle's restatement of run! (sqp_trust_region.jl:98-223) with the DEVICE QP solve behind its
    sub-optimizer hook, against the device-side driver -- same final status, objective within 1e-6, same trajectory."""
    from device_sub import attach
    eng = capi.Engine(0)
    try:
        lo, ld = [], []
        ora = attach(SqpTROracle(make(), OParams(**kw)), eng).run(lo)
        dev = SqpTR(make(), Parameters(**kw)).run(ld)
        dev.close()
        same = _prefix_equal(lo, ld, ("f", "Delta", "pinf", "mu", "inf_pr", "inf_du"), 1e-7)
        print(name, "status", ora.status, dev.status, "iters", ora.iter, dev.iter, "identical log entries", same, "of", len(lo), len(ld))
        if whole:
            assert ora.status == dev.status and ora.iter == dev.iter
            assert abs(ora.obj_val - dev.obj_val) <= 1e-6 * max(1.0, abs(ora.obj_val))
            assert np.abs(ora.x - dev.x).max() <= 1e-6 * max(1.0, np.abs(ora.x).max())
            assert same == len(lo) == len(ld)
        else:
            assert same >= 10
        if name == "case9_default":  # the reference's default parameters reach the public case9 optimum, status 0
            assert dev.status == 0 and abs(dev.obj_val - 5296.686204) <= 1e-6 * 5296.686204
    finally:
        eng.close()


@pytest.mark.parametrize("name,make,kw", [("readme_toy", ReadmeToy, dict(max_iter=100)), ("toy", ToyExample, dict(max_iter=200)),
                                          ("case9", lambda: AcopfPolar(case9()), dict(max_iter=200))])
def test_line_search_driver_parity_with_shared_subsolver(built_lib, name, make, kw):
    """BASELINE configs[1] (SQP line search on case9): the CPU restatement of the line-search driver
    (oracle/sqp_ls.py: sqp_line_search.jl:71-334) with the DEVICE QP solve behind its hook, against the device-side
    driver (host/sqp_line_search.py: penalty rule, Armijo search, merit, directional derivative, complementarity on the
    device) -- same multipliers on both sides, so the whole trajectory and the final objective (1e-6) must agree."""
    from device_sub import attach
    from oracle.sqp_ls import LsParameters as OLs, SqpLSOracle
    from sqpsolver_jl_b200.host.sqp_line_search import LsParameters, SqpLS
    eng = capi.Engine(0)
    try:
        lo, ld = [], []
        ora = attach(SqpLSOracle(make(), OLs(**kw)), eng).run(lo)
        dev = SqpLS(make(), LsParameters(**kw)).run(ld)
        dev.close()
        same = _prefix_equal(lo, ld, ("f", "phi", "alpha", "pinf", "inf_pr", "inf_du", "compl", "mu"), 1e-7)
        print(name, "LS status", ora.status, dev.status, "iters", ora.iter, dev.iter, "identical", same, "of", len(lo), len(ld),
              "obj", ora.obj_val, dev.obj_val)
        assert ora.status == dev.status and ora.iter == dev.iter
        assert same == len(lo) == len(ld)
        assert abs(ora.obj_val - dev.obj_val) <= 1e-6 * max(1.0, abs(ora.obj_val))
        assert np.abs(ora.x - dev.x).max() <= 1e-6 * max(1.0, np.abs(ora.x).max())
    finally:
        eng.close()


@pytest.mark.parametrize("key,init_mu,rtol", [("init_mu_1e5", 1e5, 1e-6), ("init_mu_1", 1.0, 1e-5)])
def test_case118_reference_iteration_limit_against_pinned_oracle_run(built_lib, key, init_mu, rtol):
    """BASELINE configs[2] with the reference's default max_iter = 3000: neither side converges (as-coded algorithm), both
    stop at the iteration limit on a feasible point -- same termination status (6) and the final objective of the pinned
    oracle run (tests/golden/case118_3000.json, 45 CPU-minutes each) to 1e-6 (bench setting) / 1e-5 (init_mu = 1: the
    stalled iterates drift along a flat valley, 5e-6 apart after 3000 nonconvex subproblems)."""
    import json
    g = json.load(open(os.path.join(HERE, "golden", "case118_3000.json")))[key]
    dev = SqpTR(AcopfPolar(synth_net(118, 186, 54, 118)), Parameters(max_iter=3000, init_mu=init_mu)).run()
    dev.close()
    print(key, "device status", dev.status, "obj", dev.obj_val, "oracle", g["status"], g["obj"])
    assert dev.status == g["status"] == 6
    assert dev.iter == g["iter"]
    assert abs(dev.obj_val - g["obj"]) <= rtol * abs(g["obj"])
    assert dev.prim_infeas <= 1e-8
