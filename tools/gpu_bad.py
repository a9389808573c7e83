"""DEVELOPMENT TOOL: replay the QPs saved by gpu_batch_diag.py on the device with verbose IPM output."""
import os, sys, pickle
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqpsolver_jl_b200 import capi
from sqpsolver_jl_b200.nlp.networks import synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
bad = pickle.load(open(sys.argv[1], 'rb'))
net = synth_net(118, 186, 54, 118)
opts = dict(warm_start=0, verbose=1, method=2)
for a in sys.argv[3:]:
    k, v = a.split('='); opts[k] = eval(v)
for t in bad[: int(sys.argv[2])]:
    nlp = AcopfPolar(net, pd=t['pd'], qd=t['qd'])
    eng = capi.Engine(0)
    eng.setup_nlp(nlp.n, nlp.m, nlp.num_linear_constraints, nlp.j_row, nlp.j_col, nlp.h_row, nlp.h_col, nlp.x_L, nlp.x_U, nlp.g_L, nlp.g_U)
    eng.set_options(**opts)
    eng.update_nlp(t['dE'], t['h_val'], t['df'], t['E'])
    p, lam, mxL, mxU, sl, st, info = eng.solve_tr(capi.PHASE_QP, t['x'], t['Delta'])
    print('b', t['b'], 'iter', t['iter'], 'status', st[0], 'ipm', info[0]['ipm_iters'], 'nf', info[0]['chol_factorizations'], flush=True)
    eng.close()
