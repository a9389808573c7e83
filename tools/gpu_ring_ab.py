"""Development aid (GPU box): ring (one resident CTA per SM, index programs streamed by cp.async.bulk) against the slot-list
launches on the case118-shaped batch.   usage: [SQPQP_PROF=1] python tools/gpu_ring_ab.py B rounds"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqpsolver_jl_b200 import capi
from sqpsolver_jl_b200.nlp.networks import synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.host.sqp_trust_region import BatchSqpTR, Parameters
SEG = ["prologue", "resid", "weights", "assemble", "factor_sparse", "assemble_slots", "factor_dense", "rhs", "fwd", "tail", "bwd",
       "ratio", "update", "epilogue", "other", "-", "ring_wait", "ring_work", "ring_bar", "ring_chunks(k)", "work_factor_sweep", "work_asm"]
B = int(sys.argv[1]); rounds = int(sys.argv[2])
capi.build()
net = synth_net(118, 186, 54, 118)
pd, qd = net.perturbed_loads(B)
cfgs = [("default", None, 0, None), ("tail 64", None, 0, dict(tail=64)), ("tail 80", None, 0, dict(tail=80)), ("tail 112", None, 0, dict(tail=112)), ("tail 128", None, 0, dict(tail=128)), ("no handoff", None, 0, dict(handoff=0)), ("handoff 32", None, 0, dict(handoff=32)), ("handoff 40", None, 0, dict(handoff=40)), ("handoff 56", None, 0, dict(handoff=56)), ("handoff 64", None, 0, dict(handoff=64)), ("default unfused", None, 0, dict(fuse=0)), ("occ1 slot lists", dict(occupancy=1), 1, None), ("occ1 ring", dict(occupancy=1), 2, None)]
if len(sys.argv) > 3:
    cfgs = [c for c in cfgs if c[0] in sys.argv[3:]]
for name, eo, ring, layout in cfgs:
    sqp = BatchSqpTR(AcopfPolar(net, pd=pd, qd=qd), B, Parameters(max_iter=rounds, init_mu=1e5), engine_options=eo, layout=layout)
    eng = sqp.optimizer.engine
    eng.set_layout(ring=ring)
    ms, its = [], []
    orig = sqp.optimizer._solve
    def hook(phase, x_k, delta, E_override=None, active=None):
        if capi.PROF_BUILD: eng.prof_read()
        out = orig(phase, x_k, delta, E_override, active)
        info = sqp.optimizer.last_info
        ms.append(eng.last_solve_ms); its.append((float(info['ipm_iters'].mean()), int(info['ipm_iters'].max())))
        if capi.PROF_BUILD:
            pr = eng.prof_read().astype(np.float64); it = info['ipm_iters']
            print("      kcycles/iter-inst: " + "  ".join(f"{s}:{v / max(it.sum(), 1) / 1e3:6.1f}" for s, v in zip(SEG, pr) if v > 0) + f"   total {pr[:15].sum() / max(it.sum(), 1) / 1e3:.1f}", flush=True)
        return out
    sqp.optimizer._solve = hook
    sqp.run()
    st = {int(k): int(v) for k, v in zip(*np.unique(sqp.status, return_counts=True))}
    shown = ms if len(ms) <= 10 else ms[:4] + ms[-4:]
    print(f"{name:18s} kernel {eng.last_solve_kernel:28s} ms/round " + " ".join(f"{m:6.1f}" for m in shown) + f"   total {sum(ms):7.1f} over {len(ms)} launches   late half {sum(ms[len(ms)//2:]):7.1f}   iters(mean,max) last {its[-1]} max-of-max {max(i[1] for i in its)}  status {st}", flush=True)
    sqp.close()
