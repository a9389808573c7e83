"""Development aid (GPU box): inertia-correction growth / decay on the WHOLE 101-round batched solve (device evaluator)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqpsolver_jl_b200.nlp.networks import synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.host.sqp_trust_region import BatchSqpTR, Parameters
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
net = synth_net(118, 186, 54, 118)
pd, qd = net.perturbed_loads(B)
for g, d in ((4.0, 3.0), (3.0, 3.0), (5.0, 2.0), (4.0, 2.0), (6.0, 3.0), (3.0, 2.0), (8.0, 3.0)):
    sqp = BatchSqpTR(AcopfPolar(net, pd=pd, qd=qd), B, Parameters(max_iter=100, init_mu=1e5), device_evaluator=True,
                     engine_options=dict(ipm_ic_growth=g, ipm_ic_decay=d))
    mx, mean = [], []
    orig = sqp.optimizer._solve
    def hook(phase, x_k, delta, E_override=None, active=None, _o=orig):
        out = _o(phase, x_k, delta, E_override, active)
        info = sqp.optimizer.last_info
        sel = np.ones(B, bool) if active is None else np.asarray(active, bool)
        mx.append(int(info['ipm_iters'][sel].max())); mean.append(float(info['ipm_iters'][sel].mean()))
        return out
    sqp.optimizer._solve = hook
    t0 = time.time(); sqp.run(); w = time.time() - t0
    st = {int(k): int(v) for k, v in zip(*np.unique(sqp.status, return_counts=True))}
    print(f"growth {g} decay {d}: kernel {sqp.optimizer.stats['solve_ms']/1e3:6.2f} s  wall {w:5.1f} s  iters mean {np.mean(mean):5.1f}  mean-of-max {np.mean(mx):5.1f}  max {max(mx)}  fallbacks {sqp.optimizer.stats['admm_iters']}  status {st}", flush=True)
    sqp.close()
