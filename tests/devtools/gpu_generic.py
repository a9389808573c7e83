"""DEVELOPMENT TOOL: generic-lane random convex QP with verbose IPM trace."""
import os, sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sqpsolver_jl_b200 import capi
from oracle import qp_solver as qs
rng = np.random.default_rng(21)
eng = capi.Engine(0)
opts = dict(verbose=1)
for a in sys.argv[1:]:
    k, v = a.split('='); opts[k] = eval(v)
eng.set_options(**opts)
for trial in range(1):
    n, m = 40, 25
    M = rng.standard_normal((n, n)); Pd = M @ M.T + 0.5 * np.eye(n)
    A = sp.random(m, n, 0.25, random_state=trial, data_rvs=rng.standard_normal).tocoo()
    q = rng.standard_normal(n) * 5; x0 = rng.uniform(-0.5, 0.5, n); Ax = A.tocsr() @ x0
    rl, ru = Ax - rng.uniform(0, 0.5, m), Ax + rng.uniform(0, 0.5, m)
    rl[:5] = ru[:5] = Ax[:5]; rl[5:8] = -np.inf
    cl, cu = np.full(n, -1.0), np.full(n, 1.0); cu[:3] = np.inf
    iu = np.triu_indices(n)
    eng.qp_setup(n, m, iu[0] + 1, iu[1] + 1, A.row + 1, A.col + 1)
    x, rd, cd, st, info = eng.qp_solve(Pd[iu], q, A.data, rl, ru, cl, cu)
    res = qs.solve_qp(sp.csr_matrix(Pd), q, A.tocsr(), rl, ru, cl, cu)
    print('status', st, 'ipm', info['ipm_iters'], 'admm', info['admm_iters'], '|dx|', np.abs(x - res.x).max(), '|drd|', np.abs(rd - res.row_dual).max(), 'rp', info['res_prim'], 'rd', info['res_dual'])
    print(qs.kkt_residuals(sp.csr_matrix(Pd), q, A.tocsr(), rl, ru, cl, cu, x, rd, cd))
    print(qs.kkt_residuals(sp.csr_matrix(Pd), q, A.tocsr(), rl, ru, cl, cu, res.x, res.row_dual, res.col_dual))
