"""Development aid (GPU box): the ~2000-bus single instance (BASELINE configs[3]) on the cooperative-grid team -- time per
QP subproblem and per interior-point iteration over the first SQP iterations, with the dense-tail size in effect."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from sqpsolver_jl_b200 import capi  # noqa: E402
from sqpsolver_jl_b200.host.sqp_trust_region import Parameters, SqpTR  # noqa: E402
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar  # noqa: E402
from sqpsolver_jl_b200.nlp.networks import synth_net  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 12
tails = [int(a) for a in sys.argv[2:]] or [-1]
for tail in tails:
    nlp = AcopfPolar(synth_net(2000, 3000, 400, 2000))
    t0 = time.time()
    d = SqpTR(nlp, Parameters(max_iter=iters, init_mu=1e5))
    eng = d.batch.optimizer.engine
    if tail >= 0:
        eng.set_layout(tail=tail)
        d.batch.optimizer.create_model(None)
    print("tail", tail, "setup %.1fs" % (time.time() - t0), "chol", eng.chol_stats(), flush=True)
    per = []
    orig = d.batch.optimizer._solve

    def hook(phase, x_k, delta, E_override=None, active=None, _o=orig, _p=per):
        r = _o(phase, x_k, delta, E_override, active)
        info = d.batch.optimizer.last_info[0]
        _p.append((phase, eng.last_solve_ms, int(info["ipm_iters"]), int(info["chol_factorizations"]), int(info["moi_status"]), int(info["admm_iters"])))
        return r

    d.batch.optimizer._solve = hook
    t0 = time.time()
    d.run()
    wall = time.time() - t0
    for p in per:
        print("   phase %d  %.1f ms  ipm %d  fact %d  status %d admm %d  -> %.2f ms/iter" % (p[0], p[1], p[2], p[3], p[4], p[5], p[1] / max(1, p[2])))
    tot = sum(p[1] for p in per)
    its = sum(p[2] for p in per)
    print("tail", tail, "kernel", eng.last_solve_kernel, "QPs", len(per), "total %.1f ms" % tot, "mean %.1f ms/QP" % (tot / len(per)),
          "%.2f ms/ipm-iteration" % (tot / max(1, its)), "wall %.1fs" % wall, "status", d.status, flush=True)
    d.close()
