"""Development aid (GPU box): run device SQP-TR trajectories on the BASELINE single-instance configs and dump status,
objective, iteration counts and the per-iteration log, for comparison with the oracle's runs."""
import collections
import pickle
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from sqpsolver_jl_b200.host.sqp_trust_region import Parameters, SqpTR  # noqa: E402
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar  # noqa: E402
from sqpsolver_jl_b200.nlp.networks import case9, synth_net  # noqa: E402

CASES = {
    "case9_default_100": (lambda: AcopfPolar(case9()), dict(max_iter=100)),
    "case9_default_3000": (lambda: AcopfPolar(case9()), dict(max_iter=3000)),
    "case118_mu1_3000": (lambda: AcopfPolar(synth_net(118, 186, 54, 118)), dict(max_iter=3000, init_mu=1.0)),
    "case118_mu1e5_3000": (lambda: AcopfPolar(synth_net(118, 186, 54, 118)), dict(max_iter=3000, init_mu=1e5)),
}
out = {}
for name in sys.argv[1:] or CASES:
    mk, kw = CASES[name]
    log = []
    t0 = time.time()
    d = SqpTR(mk(), Parameters(**kw)).run(log)
    dt = time.time() - t0
    print(name, "status", d.status, "iter", d.iter, "obj", repr(d.obj_val), "nqp", d.n_qp, "prim", d.prim_infeas, "dual", d.dual_infeas,
          "wall %.1fs" % dt, "solve_ms", d.stats["solve_ms"], flush=True)
    print("  sub_status", collections.Counter(l["sub_status"] for l in log), "accept", collections.Counter(l["accept"] for l in log))
    for l in log[-3:]:
        print("  ", {k: (f"{v:.3e}" if isinstance(v, float) else v) for k, v in l.items()})
    out[name] = {"status": d.status, "iter": d.iter, "obj": d.obj_val, "log": log}
    d.close()
pickle.dump(out, open("gpurun_out/traj_dev.pkl", "wb"))
