"""Development aid (GPU box): the two driver-parity comparisons with the ring on / off, verbose."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "support"))
from sqpsolver_jl_b200 import capi
from sqpsolver_jl_b200.host.sqp_trust_region import Parameters, SqpTR
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.nlp.networks import case9, synth_net
from oracle.sqp_tr import Parameters as OParams, SqpTROracle
from device_sub import attach
from jump_replay import JumpReplay

orig_init = capi.Engine.__init__
RING = [0]
def patched(self, device=0):
    orig_init(self, device)
    self.set_layout(ring=RING[0])
capi.Engine.__init__ = patched

for ring in (0, 1):
    RING[0] = ring
    print("==== ring", ring)
    make = lambda: AcopfPolar(synth_net(118, 186, 54, 118)); kw = dict(max_iter=12, init_mu=1e5)
    eng = capi.Engine(0)
    lo, ld = [], []
    o = SqpTROracle(make(), OParams(**kw)); ora = attach(o, eng).run(lo)
    dev = SqpTR(make(), Parameters(**kw)).run(ld); dev.close()
    print("n_qp oracle-driver", o.n_qp, "device driver", dev.n_qp if hasattr(dev, "n_qp") else None)
    for a, b in list(zip(lo, ld))[:4]:
        print({k: a[k] for k in ("fr", "f", "Delta", "pinf", "mu", "inf_pr", "inf_du")})
        print({k: b[k] for k in ("fr", "f", "Delta", "pinf", "mu", "inf_pr", "inf_du")})
    eng.close()
    make = lambda: AcopfPolar(case9()); kw = dict(max_iter=100, init_mu=1e4)
    eng = capi.Engine(0)
    replay = []
    def factory(data):
        r = JumpReplay(data, eng); replay.append(r); return r
    lg = []
    g = SqpTROracle(make(), OParams(**kw), sub_factory=factory); gen = g.run(lg)
    print("case9 generic lane: status", gen.status, "n_qp", g.n_qp, "replays", len(replay), [r.n_solves for r in replay], "setups", [r.n_setups for r in replay], "iters", gen.iter, "fr iters", sum(1 for a in lg if a["fr"]))
    eng.close()
