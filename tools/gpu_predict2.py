"""Development aid (GPU box): which quantities known BEFORE a batched solve predict its slow instances?
usage: python tools/gpu_predict2.py [B] [rounds]   -> gpurun_out/r2q_features.npz"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqpsolver_jl_b200.nlp.networks import synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.host.sqp_trust_region import BatchSqpTR, Parameters
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 60
net = synth_net(118, 186, 54, 118)
pd, qd = net.perturbed_loads(B)
sqp = BatchSqpTR(AcopfPolar(net, pd=pd, qd=qd), B, Parameters(max_iter=rounds, init_mu=1e5), device_evaluator=True)
eng = sqp.optimizer.engine
orig = sqp.optimizer._solve
rows = []
def hook(phase, x_k, delta, E_override=None, active=None):
    feats = dict(Delta=sqp.Delta.copy(), acc=sqp.step_acceptance.copy(), mu=sqp.mu.copy(), pinf=np.abs(sqp.p).max(axis=1),
                 prim=sqp.prim_infeas.copy(), dual=sqp.dual_infeas.copy(), lam=np.abs(sqp.lam).max(axis=1),
                 nact_lo=(sqp.mult_x_L > 1e-8).sum(axis=1), nact_up=(sqp.mult_x_U < -1e-8).sum(axis=1), it=sqp.iter.copy())
    out = orig(phase, x_k, delta, E_override, active)
    if phase == 0:
        info = sqp.optimizer.last_info
        feats.update(iters=info['ipm_iters'].astype(np.int64), facts=info['chol_factorizations'].astype(np.int64),
                     active=np.ones(B, bool) if active is None else np.asarray(active, bool).copy(), ms=np.array([eng.last_solve_ms]),
                     status=info['moi_status'].astype(np.int64))
        rows.append(feats)
    return out
sqp.optimizer._solve = hook
t0 = time.time(); sqp.run(); print(f"wall {time.time() - t0:.1f} s rounds {len(rows)}", flush=True)
rec = {}
for k, f in enumerate(rows):
    for name, v in f.items():
        rec[f"r{k}_{name}"] = v
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/r2q_features.npz", **rec)
sqp.close()
