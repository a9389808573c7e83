"""Development aid: one tiny interleaved solve (for compute-sanitizer)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from sqpsolver_jl_b200 import capi
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.nlp.networks import case9
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4
net = case9(); B = 5
pd, qd = net.perturbed_loads(B)
nlp = AcopfPolar(net, pd=pd, qd=qd)
x = np.broadcast_to(nlp.x0, (B, nlp.n)).copy()
df = np.empty((B, nlp.n)); nlp.eval_grad_f(x, df)
E = np.empty((B, nlp.m)); nlp.eval_g(x, E)
dE = np.empty((B, nlp.nnz_jac_coo)); nlp.eval_jac_g(x, dE)
hv = np.empty((B, nlp.nnz_hess_coo)); nlp.eval_h(x, 1.0, np.zeros((B, nlp.m)), hv)
eng = capi.Engine(0)
eng.set_layout(G=G, threads=512)
eng.setup_nlp(nlp.n, nlp.m, nlp.num_linear_constraints, nlp.j_row, nlp.j_col, nlp.h_row, nlp.h_col, nlp.x_L, nlp.x_U, nlp.g_L, nlp.g_U, batch=B)
eng.update_nlp(dE, hv, df, E)
for ph in (capi.PHASE_QP, capi.PHASE_FR, capi.PHASE_LP):
    r = eng.solve_tr(ph, x, 0.3 if ph != capi.PHASE_LP else np.inf)
    print("phase", ph, "status", r[5], "iters", r[6]["ipm_iters"], "nfact", r[6]["chol_factorizations"], "|p|", np.abs(r[0]).max(axis=1))
eng.close()
