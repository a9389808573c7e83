"""Development aid (GPU box): record reject -> re-solve pairs of the batched case118-shaped solve ("only the box changed":
sqp_trust_region.jl:134, 574-577) for the warm-start prototype tests/devtools/proto_warm.py.
usage: python tools/gpu_record_pairs.py [B] [rounds] [max_pairs]  -> gpurun_out/r2u_pairs118.pkl"""
import os, pickle, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqpsolver_jl_b200.nlp.networks import synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.host.sqp_trust_region import BatchSqpTR, Parameters
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 48
net = synth_net(118, 186, 54, 118)
pd, qd = net.perturbed_loads(B)
sqp = BatchSqpTR(AcopfPolar(net, pd=pd, qd=qd), B, Parameters(max_iter=rounds, init_mu=1e5), device_evaluator=False)
orig = sqp.optimizer._solve
pairs = []
prev_delta = np.full(B, np.nan)
nrej = [0, 0]
def hook(phase, x_k, delta, E_override=None, active=None):
    if phase == 0:
        sel = np.ones(B, bool) if active is None else np.asarray(active, bool)
        re = sel & ~sqp.step_acceptance & np.isfinite(prev_delta)
        nrej[0] += int(re.sum()); nrej[1] += int(sel.sum())
        for b in np.nonzero(re)[0]:
            if len(pairs) >= cap:
                break
            t = dict(b=int(b), iter=int(sqp.iter[b]), x=sqp.x[b].copy(), dE=sqp.dE[b].copy(), h_val=sqp.h_val[b].copy(), df=sqp.df[b].copy(),
                     E=sqp.E[b].copy(), pd=pd[b].copy(), qd=qd[b].copy())
            a = dict(t); a["Delta"] = float(prev_delta[b])
            c = dict(t); c["Delta"] = float(np.broadcast_to(delta, (B,))[b])
            pairs.append((a, c))
    out = orig(phase, x_k, delta, E_override, active)
    if phase == 0:
        sel = np.ones(B, bool) if active is None else np.asarray(active, bool)
        prev_delta[sel] = np.broadcast_to(delta, (B,))[sel]
    return out
sqp.optimizer._solve = hook
t0 = time.time(); sqp.run()
print(f"wall {time.time() - t0:.1f} s; re-solves after a rejected step: {nrej[0]} of {nrej[1]} QP-phase subproblems ({100.0 * nrej[0] / max(nrej[1], 1):.2f} %); recorded {len(pairs)} pairs", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
pickle.dump(pairs, open("gpurun_out/r2u_pairs118.pkl", "wb"))
sqp.close()
