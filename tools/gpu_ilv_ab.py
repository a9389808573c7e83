"""Development aid (GPU box): A/B of the batched interior-point launch layouts on BASELINE configs[4] (1024 perturbed-load
case118-shaped instances): one CTA per instance (ipm.cuh) against G instances interleaved per CTA (ilv.cuh).  Prints, per
layout, the solve-kernel time of the first SQP rounds (CUDA events on the engine's stream), iteration statistics and
statuses, so that profiles/r02_layout_ab.md can be written from one run."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from sqpsolver_jl_b200 import capi  # noqa: E402
from sqpsolver_jl_b200.host.sqp_trust_region import BatchSqpTR, Parameters  # noqa: E402
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar  # noqa: E402
from sqpsolver_jl_b200.nlp.networks import synth_net  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ROUNDS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
net = synth_net(118, 186, 54, seed=118)
pd, qd = net.perturbed_loads(B)
LAYOUTS = [dict(G=1), dict(G=2, threads=512, ctas_per_sm=2), dict(G=4, threads=512, ctas_per_sm=1), dict(G=4, threads=1024),
           dict(G=4, threads=512, ctas_per_sm=2), dict(G=8, threads=512), dict(G=8, threads=1024),
           dict(G=4, threads=512, ctas_per_sm=1, tail=64), dict(G=4, threads=1024, tail=64)]
if len(sys.argv) > 3:
    LAYOUTS = [eval("dict(%s)" % a) for a in sys.argv[3:]]
base = None
for lay in LAYOUTS:
    nlp = AcopfPolar(net, pd=pd, qd=qd)
    try:
        sqp = BatchSqpTR(nlp, B, Parameters(max_iter=ROUNDS, init_mu=1e5), layout=lay)
    except capi.SqpQpError as e:
        print("layout", lay, "FAILED setup:", e, flush=True)
        continue
    eng = sqp.optimizer.engine
    per = []
    orig = sqp.optimizer._solve

    def hook(phase, x_k, delta, E_override=None, active=None, _o=orig, _e=eng, _p=per):
        r = _o(phase, x_k, delta, E_override, active)
        info = sqp.optimizer.last_info
        sel = slice(None) if active is None else np.asarray(active, bool)
        _p.append((phase, _e.last_solve_ms, int(np.sum(sel)) if active is not None else B, float(info["ipm_iters"][sel].mean()),
                   int(info["ipm_iters"][sel].max()), int((info["admm_iters"][sel] > 0).sum())))
        return r

    sqp.optimizer._solve = hook
    t0 = time.time()
    sqp.run()
    wall = time.time() - t0
    qp = [p for p in per if p[0] == capi.PHASE_QP]
    ms = [p[1] for p in qp]
    st = dict(zip(*[a.tolist() for a in np.unique(sqp.status, return_counts=True)]))
    obj = sqp.obj_val.copy()
    if base is None:
        base = obj
    print("layout", lay, "chol", eng.chol_stats(), flush=True)
    print("   QP rounds ms:", " ".join("%.1f" % v for v in ms), "| total %.1f ms | other phases %.1f ms" % (sum(ms), sum(p[1] for p in per if p[0] != capi.PHASE_QP)))
    print("   ipm iters mean/max per round:", " ".join("%.1f/%d" % (p[3], p[4]) for p in qp), "| fallbacks", sum(p[5] for p in per))
    print("   status", st, "wall %.1fs" % wall, "max rel obj diff vs first layout %.2e" % float(np.max(np.abs(obj - base) / np.maximum(1.0, np.abs(base)))), flush=True)
    sqp.close()
