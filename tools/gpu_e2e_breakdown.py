"""Where the host-buffer (e2e) step of bench.py spends its time: wall clock of every C-ABI call of one SQP round on the
1024-instance case118-shaped batch against the device time of the solve kernel.  Run on the GPU box:
    python tools/gpu_e2e_breakdown.py [batch]
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from sqpsolver_jl_b200 import capi  # noqa: E402
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar  # noqa: E402
from sqpsolver_jl_b200.nlp.networks import synth_net  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
net = synth_net(118, 186, 54, seed=118)
pd, qd = net.perturbed_loads(B)
nlp = AcopfPolar(net, pd=pd, qd=qd)
rng = np.random.default_rng(3)
x = np.clip(np.broadcast_to(nlp.x0, (B, nlp.n)) + 0.02 * rng.standard_normal((B, nlp.n)), nlp.x_L, nlp.x_U)
lam = np.zeros((B, nlp.m))
df = np.empty((B, nlp.n)); nlp.eval_grad_f(x, df)
E = np.empty((B, nlp.m)); nlp.eval_g(x, E)
dE = np.empty((B, nlp.nnz_jac_coo)); nlp.eval_jac_g(x, dE)
hv = np.empty((B, nlp.nnz_hess_coo)); nlp.eval_h(x, 1.0, lam, hv)
f = np.atleast_1d(nlp.eval_f(x)).astype(float)
mu = np.full(B, 1e4)
z = np.zeros_like(x)
eng = capi.Engine()
eng.setup_nlp(nlp.n, nlp.m, nlp.num_linear_constraints, nlp.j_row, nlp.j_col, nlp.h_row, nlp.h_col, nlp.x_L, nlp.x_U, nlp.g_L,
              nlp.g_U, batch=B)
eng.set_options(warm_start=0)
import os
eng.reuse_outputs = os.environ.get("REUSE", "1") == "1"
mb = lambda *a: sum(v.nbytes for v in a) / 1e6
acc = {}
for it in range(6):
    t = [time.perf_counter()]
    eng.update_nlp(dE, hv, df, E); t.append(time.perf_counter())
    eng.merit(x, z, E, f, mu); t.append(time.perf_counter())
    eng.kt_residuals(lam, z, z); t.append(time.perf_counter())
    out = eng.solve_tr(capi.PHASE_QP, x, 10.0); t.append(time.perf_counter())
    if it >= 2:
        for k, name in enumerate(("update_nlp", "merit", "kt_residuals", "solve_tr")):
            acc.setdefault(name, []).append((t[k + 1] - t[k]) * 1e3)
        acc.setdefault("solve_kernel", []).append(eng.last_solve_ms)
print("batch", B, "H2D MB: update", round(mb(dE, hv, df, E), 1), "merit", round(mb(x, z, E, f, mu), 1), "kt", round(mb(lam, z, z), 1),
      "solve in", round(mb(x), 1), "out", round(mb(*[o for o in out[:5]]), 1))
for k, v in acc.items():
    print("%-14s %8.2f ms" % (k, float(np.median(v))))
eng.close()
