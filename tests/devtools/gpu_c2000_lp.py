"""DEVELOPMENT TOOL: 2000-bus instance, start-point projection and first QP on the device vs the CPU oracle."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sqpsolver_jl_b200 import capi
from sqpsolver_jl_b200.nlp.networks import synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.host.subproblem import QpDevice
from oracle.coo import CooMatrix
from oracle.subproblem import sub_optimize_lp
nlp = AcopfPolar(synth_net(2000, 3000, 400, 2000))
x0 = np.asarray(nlp.x0, dtype=float).copy()
qp = QpDevice(nlp, batch=1, engine_options=dict(verbose=int(os.environ.get("VERB", "0")), ipm_max_iter=int(os.environ.get("IPMIT", "200")), team=int(os.environ.get("TEAM", "0")))); qp.create_model()
X = x0[None, :]
dE = np.zeros((1, nlp.nnz_jac_coo)); nlp.eval_jac_g(X, dE)
E = np.zeros((1, nlp.m)); nlp.eval_g(X, E)
df = np.zeros((1, nlp.n)); nlp.eval_grad_f(X, df)
hv = np.zeros((1, nlp.nnz_hess_coo)); nlp.eval_h(X, 1.0, np.zeros((1, nlp.m)), hv)
qp.engine.update_nlp(dE, hv, df, E)
t0 = time.time(); p, lam, mxU, mxL, st = qp.sub_optimize_lp(x0[None, :]); td = time.time() - t0
J = CooMatrix(nlp.j_row, nlp.j_col, nlp.m, nlp.n); J.fill(dE[0])
t0 = time.time(); xo, lo, uo, lo2, so = sub_optimize_lp(J.to_scipy(), nlp.g_L, nlp.g_U, nlp.x_L, nlp.x_U, x0, nlp.num_linear_constraints, nlp.m); to = time.time() - t0
print('device status', st, 'time', td, 'oracle status', so, 'time', to)
print('max |x_dev - x_oracle|', np.abs(p[0] - xo).max(), ' |x_oracle - x0|', np.abs(xo - x0).max(), ' |x_dev - x0|', np.abs(p[0] - x0).max())
print('objective dev', ((p[0] - x0) ** 2).sum(), 'oracle', ((xo - x0) ** 2).sum())
A = J.to_scipy().tocsr()[: nlp.num_linear_constraints]
for name, x in (('dev', p[0]), ('oracle', xo)):
    ax = A @ x
    print(name, 'row viol', max(0, (nlp.g_L[: nlp.num_linear_constraints] - ax).max(), (ax - nlp.g_U[: nlp.num_linear_constraints]).max()), 'bound viol', max(0, (nlp.x_L - x).max(), (x - nlp.x_U).max()))
print('info', qp.last_info[0])
