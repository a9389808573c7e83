"""Development aid (GPU box): kernel time of the QP-phase and restoration-phase launches per SQP round of the batched solve
(are two back-to-back launches per round worth fusing?).   usage: python tools/gpu_phase_share.py B rounds"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqpsolver_jl_b200.nlp.networks import synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.host.sqp_trust_region import BatchSqpTR, Parameters
B = int(sys.argv[1]); rounds = int(sys.argv[2])
net = synth_net(118, 186, 54, 118)
pd, qd = net.perturbed_loads(1024)
sqp = BatchSqpTR(AcopfPolar(net, pd=pd[:B], qd=qd[:B]), B, Parameters(max_iter=rounds, init_mu=1e5), device_evaluator=True)
eng = sqp.optimizer.engine
orig = sqp.optimizer._solve
rows = []
def hook(phase, x_k, delta, E_override=None, active=None):
    out = orig(phase, x_k, delta, E_override, active)
    info = sqp.optimizer.last_info
    sel = np.ones(B, bool) if active is None else np.asarray(active, bool)
    rows.append((sqp.rounds, int(phase), int(sel.sum()), eng.last_solve_ms, int(info['ipm_iters'][sel].max()), float(info['ipm_iters'][sel].mean())))
    return out
sqp.optimizer._solve = hook
sqp.run()
tot = {0: 0.0, 1: 0.0, 3: 0.0}; cnt = {0: 0, 1: 0, 3: 0}
for r in rows:
    tot[r[1]] = tot.get(r[1], 0.0) + r[3]; cnt[r[1]] = cnt.get(r[1], 0) + r[2]
print(f"B {B} rounds {sqp.rounds}: QP launches {tot[0]:.1f} ms ({cnt[0]} solves), restoration launches {tot[1]:.1f} ms ({cnt[1]} solves), projection {tot[3]:.1f} ms")
both = {}
for r in rows:
    both.setdefault(r[0], {})[r[1]] = r
fused = sum(max(v[p][3] for p in v) for v in both.values()); seq = sum(sum(v[p][3] for p in v) for v in both.values())
print(f"sum over rounds of (QP + FR) {seq:.1f} ms; of max(QP, FR) {fused:.1f} ms (bound of a fused launch for a one-wave shard)")
for k in sorted(both)[:: max(1, len(both) // 12)]:
    print("  round", k, {p: (v[2], round(v[3], 2), v[4]) for p, v in both[k].items()})
sqp.close()
