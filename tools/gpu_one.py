"""DEVELOPMENT TOOL: solve selected golden QPs on the device with verbose IPM output."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqpsolver_jl_b200 import capi
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.nlp.networks import case9

g = np.load("tests/golden/case9_mu1e4.npz")
nlp = AcopfPolar(case9())
eng = capi.Engine(0)
eng.setup_nlp(nlp.n, nlp.m, nlp.num_linear_constraints, nlp.j_row, nlp.j_col, nlp.h_row, nlp.h_col, nlp.x_L, nlp.x_U, nlp.g_L, nlp.g_U)
print(eng.chol_stats())
opts = dict(warm_start=0, verbose=1, method=2, ipm_max_iter=12)
for a in sys.argv[2:]:
    k, v = a.split("=")
    opts[k] = type(getattr(eng.opts, k))(eval(v))
eng.set_options(**opts)
for k in eval(sys.argv[1]):
    eng.update_nlp(g["qp_dE"][k], g["qp_h_val"][k], g["qp_df"][k], g["qp_E"][k])
    p, lam, mxL, mxU, sl, st, info = eng.solve_tr(capi.PHASE_QP, g["qp_x"][k], g["qp_Delta"][k])
    print("k", k, "status", st[0], "ipm", info[0]["ipm_iters"], "nf", info[0]["chol_factorizations"], "|dp|",
          np.abs(p[0] - g["qp_p"][k]).max(), flush=True)
