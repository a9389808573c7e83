"""DEVELOPMENT TOOL: device line-search driver next to its CPU restatement."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sqpsolver_jl_b200.nlp.networks import case9
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.nlp.toy import ToyExample, ReadmeToy
from sqpsolver_jl_b200.host.sqp_line_search import SqpLS, LsParameters
from oracle.sqp_ls import SqpLSOracle, LsParameters as OP
for name, mk, kw in (('readme', ReadmeToy, dict(max_iter=100)), ('toy', ToyExample, dict(max_iter=100)), ('case9', lambda: AcopfPolar(case9()), dict(max_iter=int(os.environ.get('IT', '30'))))):
    dl, ol = [], []
    t0 = time.time(); d = SqpLS(mk(), LsParameters(**kw)).run(dl); td = time.time() - t0
    t0 = time.time(); o = SqpLSOracle(mk(), OP(**kw)).run(ol); to = time.time() - t0
    print(f"{name}: DEVICE status={d.status} obj={d.obj_val:.10g} iter={d.iter} nqp={d.n_qp} {td:.2f}s | ORACLE status={o.status} obj={o.obj_val:.10g} iter={o.iter} nqp={o.n_qp} {to:.2f}s | rel obj diff {abs(d.obj_val-o.obj_val)/max(1,abs(o.obj_val)):.2e} xdiff {np.abs(d.x-o.x).max():.2e}", flush=True)
    for a, b in list(zip(dl, ol))[:int(os.environ.get('SHOW', '6'))]:
        print('    ', a['iter'], 'FR' if a['fr'] else '  ', f"dev f={a['f']:.8e} phi={a['phi']:.6e} a={a['alpha']:.3e} p={a['pinf']:.3e} pr={a['inf_pr']:.2e} du={a['inf_du']:.2e} c={a['compl']:.1e} | ora f={b['f']:.8e} phi={b['phi']:.6e} a={b['alpha']:.3e} p={b['pinf']:.3e} du={b['inf_du']:.2e} c={b['compl']:.1e}")
