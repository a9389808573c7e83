"""Development aid (GPU box): one engine over the whole shard against G engines over 1/G of it each, replaying the same recorded
SQP rounds through the device-pointer API on G streams -- do the launches of the other groups fill the straggler tail of a launch?
usage: python tools/gpu_groups_ab.py B steps [G ...]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqpsolver_jl_b200 import capi
from sqpsolver_jl_b200.nlp.networks import synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.host.sqp_trust_region import BatchSqpTR, Parameters
B = int(sys.argv[1]); steps = int(sys.argv[2]); Gs = [int(a) for a in sys.argv[3:]] or [1, 2, 4]
net = synth_net(118, 186, 54, 118)
pd, qd = net.perturbed_loads(B)
nlp = AcopfPolar(net, pd=pd, qd=qd)
sqp = BatchSqpTR(nlp, B, Parameters(max_iter=40, init_mu=1e5))
SEL = {8, 14, 20, 26, 32, 38}
rec = []
orig = sqp.optimizer._solve
def hook(phase, x_k, delta, E_override=None, active=None):
    if sqp.rounds in SEL and phase in (capi.PHASE_QP, capi.PHASE_MIXED):
        aq = np.asarray(active[0] if phase == capi.PHASE_MIXED else (np.ones(B) if active is None else active), np.int32)
        af = np.asarray(active[1], np.int32) if phase == capi.PHASE_MIXED else np.zeros(B, np.int32)
        rec.append({k: v.copy() for k, v in (("dE", sqp.dE), ("h_val", sqp.h_val), ("df", sqp.df), ("E", sqp.E), ("x", sqp.x), ("Delta", sqp.Delta))} | {"qp": aq.copy(), "fr": af.copy()})
    return orig(phase, x_k, delta, E_override, active)
sqp.optimizer._solve = hook
sqp.run(); sqp.close()
R = len(rec)
print(f"recorded {R} rounds; instances in restoration per round: {[int(r['fr'].sum()) for r in rec]}", flush=True)
dev = torch.device("cuda", 0)
for G in Gs:
    engs, data = [], []
    for g in range(G):
        lo, hi = g * B // G, (g + 1) * B // G
        e = capi.Engine(0)
        sub = AcopfPolar(net, pd=pd[lo:hi], qd=qd[lo:hi])
        e.setup_nlp(sub.n, sub.m, sub.num_linear_constraints, sub.j_row, sub.j_col, sub.h_row, sub.h_col, sub.x_L, sub.x_U, sub.g_L, sub.g_U, batch=hi - lo)
        engs.append(e)
        data.append([{k: torch.from_numpy(np.ascontiguousarray(r[k][lo:hi])).to(dev) for k in r} for r in rec])
    def step(g, j):
        e, d = engs[g], data[g][j % R]
        e.update_nlp_device(d["dE"].data_ptr(), d["h_val"].data_ptr(), d["df"].data_ptr(), d["E"].data_ptr())
        if int(rec[j % R]["fr"][g * B // G:(g + 1) * B // G].sum()):
            e.solve_tr_mixed_device(d["x"].data_ptr(), d["Delta"].data_ptr(), d["qp"].data_ptr(), d["fr"].data_ptr())
        else:
            e.solve_tr_device(capi.PHASE_QP, d["x"].data_ptr(), d["Delta"].data_ptr(), None, d["qp"].data_ptr())
    for j in range(2):
        for g in range(G): step(g, j)
    torch.cuda.synchronize()
    for e in engs: e.sync()
    t0 = time.perf_counter()
    for j in range(steps):
        for g in range(G): step(g, j)
    for e in engs: e.sync()
    dt = time.perf_counter() - t0
    units = sum(int(rec[j % R]["qp"].sum() + rec[j % R]["fr"].sum()) for j in range(steps))
    print(f"G = {G}: {steps} steps in {dt * 1e3:8.1f} ms = {dt * 1e3 / steps:6.2f} ms per step, {units / dt:8.0f} QP solves/s   kernel {engs[0].last_solve_kernel}", flush=True)
    for e in engs: e.close()
    del data
    torch.cuda.empty_cache()
