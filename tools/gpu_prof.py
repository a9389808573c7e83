"""DEVELOPMENT TOOL: in-kernel phase profile of the batched solve (build with SQPQP_PROF=1).

usage: SQPQP_PROF=1 python tools/gpu_prof.py B iters [opt=value ...]
"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqpsolver_jl_b200 import capi
from sqpsolver_jl_b200.nlp.networks import synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.host.sqp_trust_region import BatchSqpTR, Parameters
SEG = ["prologue", "resid", "weights", "assemble", "factor_sparse", "assemble_slots", "factor_dense", "rhs", "fwd", "tail", "bwd",
       "ratio", "update", "epilogue", "other"]
B = int(sys.argv[1]); iters = int(sys.argv[2])
eo = {}
for a in sys.argv[3:]:
    k, v = a.split('='); eo[k] = eval(v)
capi.build()
net = synth_net(118, 186, 54, 118)
pd, qd = net.perturbed_loads(B)
sqp = BatchSqpTR(AcopfPolar(net, pd=pd, qd=qd), B, Parameters(max_iter=iters, init_mu=1e5), engine_options=eo or None)
eng = sqp.optimizer.engine
print(eng.chol_stats())
orig = sqp.optimizer._solve
def hook(phase, x_k, delta, E_override=None, active=None):
    eng.prof_read()
    out = orig(phase, x_k, delta, E_override, active)
    pr = eng.prof_read().astype(np.float64)
    info = sqp.optimizer.last_info
    it = info['ipm_iters']; nf = info['chol_factorizations']
    tot = pr.sum()
    print(f"round {sqp.rounds:3d} phase {phase} ms {eng.last_solve_ms:8.1f} ipm it mean {it.mean():5.1f} max {it.max():3d} nfact mean {nf.mean():5.1f} max {nf.max():3d}  cta-cycles/iter {tot / max(it.sum(), 1) / 1e3:8.1f}k", flush=True)
    if tot > 0:
        print("      " + "  ".join(f"{s}:{100 * v / tot:4.1f}%" for s, v in zip(SEG, pr) if v > 0), flush=True)
        print("      kcycles per iteration-instance: " + "  ".join(f"{s}:{v / max(it.sum(), 1) / 1e3:6.1f}" for s, v in zip(SEG, pr) if v > 0), flush=True)
    return out
sqp.optimizer._solve = hook
t0 = time.time(); sqp.run(); print('total', time.time() - t0, 'timers', sqp.timers)
