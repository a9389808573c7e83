"""Unit checks of the oracle's building blocks (CPU only)."""
import numpy as np
import scipy.sparse as sp
from scipy.optimize import linprog

from oracle import qp_solver as qs
from oracle.coo import CooMatrix, SymCooMatrix, csc_pattern, ordered_scatter
from oracle.sqp_tr import KT_residuals, isapprox, norm_violations
from oracle.subproblem import trust_region_box
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.nlp.networks import case9, synth_net
from sqpsolver_jl_b200.nlp.toy import ToyExample


def _julia_like_scatter(rows1, cols1, vals, shape, symmetric=False):
    """Literal restatement of sqp.jl:92-117 with a dict-of-slots (sequential fp64 adds)."""
    A = {}
    for k in range(len(rows1)):
        r, c = int(rows1[k]) - 1, int(cols1[k]) - 1
        A[(r, c)] = A.get((r, c), 0.0) + vals[k]
        if symmetric and r != c:
            A[(c, r)] = A.get((c, r), 0.0) + vals[k]
    return A


def test_scatter_is_sequential_and_matches_literal_loop():
    rng = np.random.default_rng(3)
    m, n, nnz = 7, 5, 60  # heavy duplication
    r = rng.integers(1, m + 1, nnz)
    c = rng.integers(1, n + 1, nnz)
    v = rng.standard_normal(nnz) * 10.0 ** rng.integers(-8, 8, nnz)
    M = CooMatrix(r, c, m, n)
    M.fill(v)
    ref = _julia_like_scatter(r, c, v, (m, n))
    S = M.to_scipy().tocoo()
    assert len(ref) == M.nnz
    for i, j, x in zip(S.row, S.col, S.data):
        assert x == ref[(i, j)]  # bit-exact


def test_symmetric_scatter_mirrors_offdiagonals_once_per_entry():
    rng = np.random.default_rng(4)
    n, nnz = 6, 50
    r = rng.integers(1, n + 1, nnz)
    c = rng.integers(1, n + 1, nnz)  # both triangles and diagonal, duplicates
    v = rng.standard_normal(nnz)
    H = SymCooMatrix(r, c, n)
    H.fill(v)
    ref = _julia_like_scatter(r, c, v, (n, n), symmetric=True)
    S = H.to_scipy().tocoo()
    assert len(ref) == H.nnz
    for i, j, x in zip(S.row, S.col, S.data):
        assert x == ref[(i, j)]
    D = H.to_scipy().toarray()
    assert np.array_equal(D, D.T)


def test_csc_pattern_equals_scipy_csc():
    nlp = AcopfPolar(case9())
    cp, ri, slot = csc_pattern(nlp.j_row, nlp.j_col, nlp.n)
    A = sp.coo_matrix((np.ones(nlp.nnz_jac_coo), (nlp.j_row - 1, nlp.j_col - 1)), shape=(nlp.m, nlp.n)).tocsc()
    A.sum_duplicates()
    A.sort_indices()
    assert np.array_equal(cp, A.indptr) and np.array_equal(ri, A.indices)


def test_norm_violations_and_kt_residuals():
    E = np.array([1.0, -3.0, 0.5])
    gL = np.array([0.0, -2.0, 0.5])
    gU = np.array([0.5, np.inf, 0.5])
    x = np.array([2.0, -1.0])
    xL = np.array([-np.inf, 0.0])
    xU = np.array([1.0, 5.0])
    assert norm_violations(E, gL, gU, x, xL, xU, 1) == 0.5 + 1.0 + 0.0 + 1.0 + 1.0
    assert norm_violations(E, gL, gU, x, xL, xU, np.inf) == 1.0
    J = sp.csr_matrix(np.array([[1.0, 2.0], [0.0, -1.0], [3.0, 0.0]]))
    df = np.array([1.0, -4.0])
    lam = np.array([0.5, 2.0, -1.0])
    mU = np.array([0.0, -0.25])
    mL = np.array([0.5, 0.0])
    num = np.max(np.abs(df + J.T @ lam + mU - mL))
    den = max(1.0, 4.0, 0.25, 0.5, 0.5 * np.sqrt(5.0), 2.0 * 1.0, 1.0 * 3.0)
    assert KT_residuals(df, lam, mU, mL, J) == num / den
    assert isapprox(1.0, 1.0 + 1e-9) and not isapprox(1.0, 1.0 + 1e-7)


def test_trust_region_box_rule_including_fallback():
    lb, ub = trust_region_box(np.array([-5.0, 0.5, -np.inf]), np.array([5.0, 2.0, -0.5]), 1.0)
    # second var: x_k below its lower bound by 0.5 -> [0.5, 1]; third: x_k above upper bound by 0.5
    assert np.array_equal(lb, [-1.0, 0.5, -1.0]) and np.array_equal(ub, [1.0, 1.0, -0.5])
    lb, ub = trust_region_box(np.array([3.0]), np.array([4.0]), 1.0)  # lb > ub -> fallback (:441-444)
    assert lb[0] == 0.0 and ub[0] == 1.0


def test_ipm_matches_highs_on_lps():
    rng = np.random.default_rng(7)
    for trial in range(4):
        n, m = 12, 8
        A = rng.standard_normal((m, n))
        x0 = rng.uniform(-1, 1, n)
        b = A @ x0
        c = rng.standard_normal(n)
        rl = b - rng.uniform(0, 1, m)
        ru = b + rng.uniform(0, 1, m)
        rl[:2] = ru[:2] = b[:2]
        res = qs.solve_qp(None, c, sp.csr_matrix(A), rl, ru, np.full(n, -2.0), np.full(n, 2.0))
        lp = linprog(c, A_ub=np.vstack([A[2:], -A[2:]]), b_ub=np.concatenate([ru[2:], -rl[2:]]), A_eq=A[:2], b_eq=b[:2],
                     bounds=[(-2, 2)] * n, method="highs")
        assert res.status == "LOCALLY_SOLVED" and lp.status == 0
        assert abs(res.obj - lp.fun) <= 1e-8 * max(1.0, abs(lp.fun))


def test_ipm_kkt_on_convex_and_nonconvex_qps():
    rng = np.random.default_rng(8)
    n, m = 25, 15
    M = rng.standard_normal((n, n))
    for P in (sp.csr_matrix(M @ M.T), sp.csr_matrix(0.5 * (M + M.T))):
        q = rng.standard_normal(n)
        A = sp.random(m, n, 0.3, random_state=1, data_rvs=rng.standard_normal).tocsr()
        x0 = rng.uniform(-0.5, 0.5, n)
        Ax = A @ x0
        rl, ru = Ax - 0.3, Ax + 0.3
        rl[:4] = ru[:4] = Ax[:4]
        xl, xu = np.full(n, -1.0), np.full(n, 1.0)
        res = qs.solve_qp(P, q, A, rl, ru, xl, xu)
        assert res.status == "LOCALLY_SOLVED"
        k = qs.kkt_residuals(P, q, A, rl, ru, xl, xu, res.x, res.row_dual, res.col_dual)
        assert k["stationarity"] < 1e-8 and k["primal"] < 1e-9 and k["complementarity"] < 1e-7


def test_ipm_infeasible_and_degenerate_rows():
    A = sp.csr_matrix(np.array([[0.0, 0.0], [1.0, 0.0]]))
    res = qs.solve_qp(None, np.array([1.0, 0.0]), A, [1.0, -1.0], [1.0, 1.0], [-1, -1], [1, 1])
    assert res.status == "INFEASIBLE" and not res.x.any()
    # zero row that is satisfied (0 in [0, inf)) must not break the interior point method
    res = qs.solve_qp(None, np.array([1.0, 1.0]), A, [0.0, -1.0], [np.inf, 1.0], [-1, -1], [1, 1])
    assert res.status == "LOCALLY_SOLVED" and np.allclose(res.x, [-1, -1], atol=1e-8)


def test_nlp_derivatives_finite_difference():
    for nlp in (ToyExample(), AcopfPolar(case9()), AcopfPolar(synth_net(14, 20, 5, seed=3))):
        rng = np.random.default_rng(0)
        x = nlp.x0 + 0.1 * rng.standard_normal(nlp.n)
        lam = rng.standard_normal(nlp.m)
        h = 1e-6
        J = nlp.dense_jacobian(x)
        Jfd = np.zeros_like(J)
        for j in range(nlp.n):
            e = np.zeros(nlp.n); e[j] = h
            gp = np.empty(nlp.m); gm = np.empty(nlp.m)
            nlp.eval_g(x + e, gp); nlp.eval_g(x - e, gm)
            Jfd[:, j] = (gp - gm) / (2 * h)
        assert np.abs(J - Jfd).max() < 1e-6 * max(1.0, np.abs(J).max())

        def gradL(z):
            g = np.empty(nlp.n); nlp.eval_grad_f(z, g)
            return g + nlp.dense_jacobian(z).T @ lam

        H = nlp.dense_hessian(x, 1.0, lam)
        Hfd = np.zeros_like(H)
        for j in range(nlp.n):
            e = np.zeros(nlp.n); e[j] = h
            Hfd[:, j] = (gradL(x + e) - gradL(x - e)) / (2 * h)
        assert np.abs(H - Hfd).max() < 1e-5 * max(1.0, np.abs(H).max())


def test_batched_callbacks_match_single():
    net = synth_net(14, 20, 5, seed=3)
    pd, qd = net.perturbed_loads(3)
    nb = AcopfPolar(net, pd=pd, qd=qd)
    assert nb.g_L.shape == (3, nb.m)
    rng = np.random.default_rng(1)
    X = nb.x0 + 0.05 * rng.standard_normal((3, nb.n))
    L = rng.standard_normal((3, nb.m))
    G = np.empty((3, nb.m)); V = np.empty((3, nb.nnz_jac_coo)); Hh = np.empty((3, nb.nnz_hess_coo))
    nb.eval_g(X, G); nb.eval_jac_g(X, V); nb.eval_h(X, 1.0, L, Hh)
    for b in range(3):
        one = AcopfPolar(net, pd=pd[b], qd=qd[b])
        g = np.empty(nb.m); v = np.empty(nb.nnz_jac_coo); hh = np.empty(nb.nnz_hess_coo)
        one.eval_g(X[b], g); one.eval_jac_g(X[b], v); one.eval_h(X[b], 1.0, L[b], hh)
        assert np.array_equal(G[b], g) and np.array_equal(V[b], v) and np.array_equal(Hh[b], hh)
        assert np.array_equal(one.g_L, nb.g_L[b])
    # Philox streams: instance b is reproducible on its own
    pd2, _ = net.perturbed_loads(2)
    assert np.array_equal(pd2, pd[:2])


def test_line_search_oracle_known_answers():
    """CPU restatement of the (uncompiled, stale) line-search driver: README toy -> x = -1, toy -> (-1, -1), status 0."""
    from oracle.sqp_ls import LsParameters, SqpLSOracle
    from sqpsolver_jl_b200.nlp.toy import ReadmeToy, ToyExample
    r = SqpLSOracle(ReadmeToy(), LsParameters(max_iter=100)).run()
    assert r.status == 0 and abs(r.x[0] + 1.0) <= 1e-8 and abs(r.obj_val) <= 1e-12
    r = SqpLSOracle(ToyExample(), LsParameters(max_iter=200)).run()
    assert r.status == 0 and np.allclose(r.x, [-1.0, -1.0], atol=1e-6)
