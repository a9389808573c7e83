"""Development aid (GPU box): B case118-shaped instances, a few SQP rounds, chosen launch -- the target of ncu captures.
usage: python tools/gpu_ring_one.py B rounds [ring=0|1|2] [occupancy=N]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqpsolver_jl_b200 import capi
from sqpsolver_jl_b200.nlp.networks import synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.host.sqp_trust_region import BatchSqpTR, Parameters
B = int(sys.argv[1]); rounds = int(sys.argv[2])
kw = dict(a.split("=") for a in sys.argv[3:])
net = synth_net(118, 186, 54, 118)
pd, qd = net.perturbed_loads(B)
eo = {"occupancy": int(kw["occupancy"])} if "occupancy" in kw else None
sqp = BatchSqpTR(AcopfPolar(net, pd=pd, qd=qd), B, Parameters(max_iter=rounds, init_mu=1e5), engine_options=eo)
eng = sqp.optimizer.engine
eng.set_layout(ring=int(kw.get("ring", 0)), handoff=(int(kw["handoff"]) if "handoff" in kw else None))
orig = sqp.optimizer._solve
def hook(phase, x_k, delta, E_override=None, active=None):  # algorithmic bytes of every QP-phase launch (bench.py's model)
    out = orig(phase, x_k, delta, E_override, active)
    if phase == capi.PHASE_QP:
        import bench
        info = sqp.optimizer.last_info
        sel = np.ones(B, bool) if active is None else np.asarray(active, bool)
        nlp = sqp.problem
        nnzJ, nnzH = int(eng.get_csr(0)[1].shape[0]), int(eng.get_csr(2)[1].shape[0])
        alg = bench.algorithmic_bytes(info, sel, nlp.n, nlp.m, nnzJ, nnzH, eng.chol_stats())
        print(f"round {sqp.rounds}: QP-phase launch, {int(sel.sum())} instances, iterations mean {info['ipm_iters'][sel].mean():.1f}, "
              f"factorisations mean {info['chol_factorizations'][sel].mean():.1f}, algorithmic bytes {alg:.6e}, kernel {eng.last_solve_kernel}", flush=True)
    return out
sqp.optimizer._solve = hook
sqp.run()
print(eng.last_solve_kernel, "solve ms total", sqp.optimizer.stats["solve_ms"], "rounds", sqp.rounds)
sqp.close()
