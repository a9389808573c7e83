"""Regenerate tests/golden/case118_3000.json (run from the repo root; ~45 min of CPU per run):
    python tests/golden/make_case118_3000.py 1      # init_mu = 1  (reference default)
    python tests/golden/make_case118_3000.py 1e5    # init_mu = 1e5 (bench.py's setting)
Prints status / iterations / objective / residuals of the oracle SQP-TR run with max_iter = 3000."""
import collections
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.sqp_tr import Parameters, SqpTROracle  # noqa: E402
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar  # noqa: E402
from sqpsolver_jl_b200.nlp.networks import synth_net  # noqa: E402

mu = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
log = []
t = time.time()
o = SqpTROracle(AcopfPolar(synth_net(118, 186, 54, 118)), Parameters(max_iter=3000, init_mu=mu)).run(log)
print("init_mu", mu, "status", o.status, "iter", o.iter, "obj", repr(o.obj_val), "prim", o.prim_infeas, "dual", o.dual_infeas,
      "%.0fs" % (time.time() - t))
print(collections.Counter(l["sub_status"] for l in log), collections.Counter(l["accept"] for l in log))
