"""DEVELOPMENT TOOL: are the slow instances of one SQP round the slow ones of the next? (for longest-first CTA ordering)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqpsolver_jl_b200.nlp.networks import synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.host.sqp_trust_region import BatchSqpTR, Parameters
B = 1024
net = synth_net(118, 186, 54, 118)
pd, qd = net.perturbed_loads(B)
sqp = BatchSqpTR(AcopfPolar(net, pd=pd, qd=qd), B, Parameters(max_iter=10, init_mu=1e5))
its = []
orig = sqp.optimizer._solve
def hook(phase, x_k, delta, E_override=None, active=None):
    out = orig(phase, x_k, delta, E_override, active)
    if phase == 0: its.append(sqp.optimizer.last_info['chol_factorizations'].copy())
    return out
sqp.optimizer._solve = hook
sqp.run()
for a, b in zip(its[:-1], its[1:]):
    top = np.argsort(-b)[:30]
    ra = np.argsort(np.argsort(-a))
    print("corr %.3f   of the 30 slowest of the next round, %d were among the 100 slowest / %d among the 300 slowest of this round; max next %d" % (np.corrcoef(a, b)[0, 1], int((ra[top] < 100).sum()), int((ra[top] < 300).sum()), b.max()))
