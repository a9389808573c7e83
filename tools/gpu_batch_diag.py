"""DEVELOPMENT TOOL: per-round solver statistics of a batched device SQP run."""
import os, sys, time, pickle
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqpsolver_jl_b200 import capi
from sqpsolver_jl_b200.nlp.networks import synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.host.sqp_trust_region import BatchSqpTR, Parameters
B = int(sys.argv[1]); iters = int(sys.argv[2])
eo = {}
for a in sys.argv[3:]:
    k, v = a.split('='); eo[k] = eval(v)
net = synth_net(118, 186, 54, 118)
pd, qd = net.perturbed_loads(B)
sqp = BatchSqpTR(AcopfPolar(net, pd=pd, qd=qd), B, Parameters(max_iter=iters, init_mu=1e5), engine_options=eo or None)
print(sqp.optimizer.engine.chol_stats())
orig = sqp.optimizer._solve
bad = []
def hook(phase, x_k, delta, E_override=None, active=None):
    t0 = time.time(); out = orig(phase, x_k, delta, E_override, active); dt = time.time() - t0
    info = sqp.optimizer.last_info; act = np.ones(B, bool) if active is None else np.asarray(active, bool)
    i = info[act]
    fb = int((i['admm_iters'] > 0).sum())
    print(f"round {sqp.rounds:3d} phase {phase} active {int(act.sum()):4d} ms {sqp.optimizer.engine.last_solve_ms:8.1f} ipm it mean {i['ipm_iters'].mean():5.1f} max {i['ipm_iters'].max():3d} nfact max {i['chol_factorizations'].max():3d} fallbacks {fb} status {dict(zip(*np.unique(i['moi_status'], return_counts=True)))}", flush=True)
    slow = int(np.argmax(np.where(act, info['ipm_iters'], -1)))
    if (fb or info['ipm_iters'][slow] >= 60) and len(bad) < 6:
        b = int(np.nonzero(act & (info['admm_iters'] > 0))[0][0]) if fb else slow
        bad.append({'b': b, 'iter': int(sqp.iter[b]), 'fr': phase == 1, 'x': sqp.x[b].copy(), 'Delta': float(sqp.Delta[b]), 'dE': sqp.dE[b].copy(), 'h_val': sqp.h_val[b].copy(), 'df': sqp.df[b].copy(), 'E': sqp.E[b].copy(), 'pd': pd[b].copy(), 'qd': qd[b].copy(), 'status': int(info['moi_status'][b])})
    return out
sqp.optimizer._solve = hook
t0 = time.time(); sqp.run(); print('total', time.time() - t0, 'status', dict(zip(*np.unique(sqp.status, return_counts=True))), 'timers', sqp.timers)
pickle.dump(bad, open('gpurun_out/bad_batch.pkl', 'wb'))
