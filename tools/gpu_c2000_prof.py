"""Development aid (GPU box, SQPQP_PROF=1 build): phase profile of the cooperative-grid team on the ~2000-bus instance."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqpsolver_jl_b200 import capi
from sqpsolver_jl_b200.host.sqp_trust_region import Parameters, SqpTR
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.nlp.networks import synth_net
SEG = ["prologue", "resid", "weights", "assemble", "factor_sparse", "assemble_slots", "factor_dense", "rhs", "fwd", "tail", "bwd",
       "ratio", "update", "epilogue", "other"]
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nlp = AcopfPolar(synth_net(2000, 3000, 400, 2000))
d = SqpTR(nlp, Parameters(max_iter=iters, init_mu=1e5))
eng = d.batch.optimizer.engine
print("chol", eng.chol_stats(), flush=True)
orig = d.batch.optimizer._solve
def hook(phase, x_k, delta, E_override=None, active=None):
    eng.prof_read()
    r = orig(phase, x_k, delta, E_override, active)
    pr = eng.prof_read().astype(np.float64)[:15]
    info = d.batch.optimizer.last_info[0]
    it = max(1, int(info["ipm_iters"]))
    nb = 148.0  # cooperative CTAs (one 256-thread CTA per SM: 255 registers)
    print(f"phase {phase} {eng.last_solve_ms:7.1f} ms ipm {it} nfact {int(info['chol_factorizations'])}  kcycles/iteration (per CTA): "
          + "  ".join(f"{s}:{v / nb / it / 1e3:6.1f}" for s, v in zip(SEG, pr) if v > 0) + f"  total {pr.sum() / nb / it / 1e3:.0f}", flush=True)
    return r
d.batch.optimizer._solve = hook
d.run(); d.close()
