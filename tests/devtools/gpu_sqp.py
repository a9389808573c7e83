"""DEVELOPMENT TOOL: run the device-backed SQP-TR driver next to the oracle (under gpurun)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sqpsolver_jl_b200.nlp.networks import case9, synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.nlp.toy import ToyExample, ReadmeToy
from sqpsolver_jl_b200.host.sqp_trust_region import SqpTR, Parameters
from oracle.sqp_tr import SqpTROracle, Parameters as OP

cases = {
    'toy': (lambda: ToyExample(), dict(max_iter=100)),
    'readme': (lambda: ReadmeToy(), dict(max_iter=100)),
    'case9': (lambda: AcopfPolar(case9()), dict(max_iter=100, init_mu=1e4)),
    'case9_default': (lambda: AcopfPolar(case9()), dict(max_iter=100)),
    'case9_soc': (lambda: AcopfPolar(case9()), dict(max_iter=100, init_mu=1e4, use_soc=True)),
    'c118': (lambda: AcopfPolar(synth_net(118, 186, 54, 118)), dict(max_iter=100, init_mu=1e5)),
    'c2000': (lambda: AcopfPolar(synth_net(2000, 3000, 400, 2000)), dict(max_iter=int(os.environ.get('C2000_ITERS', '3')), init_mu=1e5)),
}
if __name__ == "__main__":
    names = sys.argv[1].split(',') if len(sys.argv) > 1 else ['toy', 'readme', 'case9']
    run_oracle = '--no-oracle' not in sys.argv
    for name in names:
        mk, kw = cases[name]
        nlp = mk()
        log = []
        t0 = time.time(); d = SqpTR(nlp, Parameters(**kw)).run(log); td = time.time() - t0
        print(f"{name}: DEVICE status={d.status} obj={d.obj_val:.10g} iters={d.iter} nqp={d.n_qp} wall={td:.2f}s solve_ms={d.stats['solve_ms']:.1f} admm={d.stats['admm_iters']} cg={d.stats['cg_iters']} pcg={d.stats['polish_cg_iters']} polished={d.stats['polished']}/{d.stats['instance_solves']} timers={ {k: round(v,3) for k,v in d.timers.items()} }", flush=True)
        if '-v' in sys.argv:
            for l in log: print('   ', l['iter'], 'FR' if l['fr'] else '  ', 'a' if l['accept'] else 'r', f"f={l['f']:.6e} mu={l['mu']:.2e} D={l['Delta']:.2e} p={l['pinf']:.2e} pr={l['inf_pr']:.2e} du={l['inf_du']:.2e} st={l['sub_status']}")
        if run_oracle:
            t0 = time.time(); o = SqpTROracle(mk(), OP(**kw)).run(); to = time.time() - t0
            print(f"{name}: ORACLE status={o.status} obj={o.obj_val:.10g} iters={o.iter} nqp={o.n_qp} wall={to:.2f}s  | rel obj diff {abs(d.obj_val-o.obj_val)/max(1,abs(o.obj_val)):.2e} x diff {np.abs(d.x-o.x).max():.2e}", flush=True)
