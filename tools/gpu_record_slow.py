"""Development aid (GPU box): record the NLP data of the slowest and of some typical subproblems of the batched case118-shaped
solve (for the numpy prototypes of the interior-point policies).   usage: python tools/gpu_record_slow.py [B] [rounds]
-> gpurun_out/r2r_slow_qps.npz"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqpsolver_jl_b200.nlp.networks import synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.host.sqp_trust_region import BatchSqpTR, Parameters
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 22
SEL = set(int(a) for a in os.environ.get('SEL', '7,13,20').split(','))
FLIPS = int(os.environ.get('FLIPS', '-1'))
net = synth_net(118, 186, 54, 118)
pd, qd = net.perturbed_loads(B)
sqp = BatchSqpTR(AcopfPolar(net, pd=pd, qd=qd), B, Parameters(max_iter=rounds, init_mu=1e5), device_evaluator=False,
                 engine_options=(dict(ipm_ic_flips=FLIPS) if FLIPS >= 0 else None))
orig = sqp.optimizer._solve
rec = {"pd": pd, "qd": qd}
count = [0]
rng = np.random.default_rng(0)
def hook(phase, x_k, delta, E_override=None, active=None):
    out = orig(phase, x_k, delta, E_override, active)
    if phase != 0:
        return out
    count[0] += 1
    r = count[0]
    if r in SEL:
        info = sqp.optimizer.last_info
        it = info['ipm_iters'].astype(np.int64)
        sel = np.ones(B, bool) if active is None else np.asarray(active, bool)
        slow = np.argsort(-(it * sel))[:8]
        typ = rng.choice(np.nonzero(sel & (it <= np.median(it[sel])))[0], 4, replace=False)
        ids = np.concatenate([slow, typ])
        rec[f"r{r}_ids"] = ids
        for name, arr in (("x", sqp.x), ("Delta", sqp.Delta), ("dE", sqp.dE), ("h_val", sqp.h_val), ("df", sqp.df), ("E", sqp.E), ("p", out[0])):
            rec[f"r{r}_{name}"] = np.asarray(arr)[ids].copy()
        rec[f"r{r}_iters"] = it[ids]; rec[f"r{r}_facts"] = info['chol_factorizations'][ids].astype(np.int64)
        rec[f"r{r}_status"] = info['moi_status'][ids].astype(np.int64); rec[f"r{r}_obj"] = info['objective'][ids].copy()
        print(f"round {r}: recorded {ids.tolist()} iters {it[ids].tolist()} facts {info['chol_factorizations'][ids].tolist()}", flush=True)
    return out
sqp.optimizer._solve = hook
t0 = time.time(); sqp.run(); print(f"wall {time.time() - t0:.1f} s", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed(os.environ.get("OUT", "gpurun_out/r2r_slow_qps.npz"), **rec)
sqp.close()
