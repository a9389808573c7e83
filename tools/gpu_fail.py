"""DEVELOPMENT TOOL: run device SQP with trace, dump non-OK QPs for offline analysis."""
import sys, os, pickle
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.gpu_sqp import cases
from sqpsolver_jl_b200.host.sqp_trust_region import SqpTR, Parameters
from sqpsolver_jl_b200 import capi
name = sys.argv[1]
eo = {}
for a in sys.argv[2:]:
    k, v = a.split('='); eo[k] = eval(v)
mk, kw = cases[name]
trace = []
d = SqpTR(mk(), Parameters(**kw), engine_options=eo or None).run(trace=trace)
print(name, 'status', d.status, 'obj', d.obj_val, 'iters', d.iter, 'solve_ms', round(d.stats['solve_ms'], 1))
bad = []
for t in trace:
    i = t['info']
    flag = t['status'] not in (4, 5)
    print(f"it{t['iter']:3d} {'FR' if t['fr'] else 'QP'} D={t['Delta']:.2e} st={capi.MOI_NAMES.get(t['status'], t['status'])} admm={i['admm_iters']} cg={i['cg_iters']} ipm={i['ipm_iters']} nf={i['chol_factorizations']} ptry={i['polish_tries']} pcg={i['polish_cg_iters']} pol={i['polished']} rho={i['rho']:.1e} rbf={i['rho_box_floor']:.2e} rp={i['res_prim']:.1e} rd={i['res_dual']:.1e}" + ('  <<<<' if flag else ''))
    if flag:
        bad.append({k: v for k, v in t.items() if k != 'info'})
pickle.dump(bad, open(f'gpurun_out/bad_{name}.pkl', 'wb'))
