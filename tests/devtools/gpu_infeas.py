"""DEVELOPMENT TOOL: device status on the oracle-infeasible QPs of the case9_default golden trajectory."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sqpsolver_jl_b200 import capi
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.nlp.networks import case9
g = np.load("tests/golden/case9_default.npz"); nlp = AcopfPolar(case9())
eng = capi.Engine(0)
eng.setup_nlp(nlp.n, nlp.m, nlp.num_linear_constraints, nlp.j_row, nlp.j_col, nlp.h_row, nlp.h_col, nlp.x_L, nlp.x_U, nlp.g_L, nlp.g_U)
eng.set_options(warm_start=0)
for k in range(len(g["qp_status"])):
    if int(g["qp_status"][k]) not in (2, 5):
        continue
    eng.set_options(verbose=0)
    eng.update_nlp(g["qp_dE"][k], g["qp_h_val"][k], g["qp_df"][k], g["qp_E"][k])
    out = eng.solve_tr(capi.PHASE_QP, g["qp_x"][k], g["qp_Delta"][k])
    i = out[6][0]
    print("k", k, "Delta", g["qp_Delta"][k], "dev status", out[5][0], "ipm", i["ipm_iters"], "admm", i["admm_iters"], "rp", i["res_prim"], flush=True)
    if out[5][0] not in (2, 5) and "-v" in sys.argv:
        eng.set_options(verbose=1, method=2)
        eng.solve_tr(capi.PHASE_QP, g["qp_x"][k], g["qp_Delta"][k])
        eng.set_options(verbose=0, method=0)
