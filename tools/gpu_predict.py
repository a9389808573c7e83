"""Development aid (GPU box): can the stragglers of a batched solve be predicted from the first interior-point iterations, and
what would starting them first be worth?   usage: python tools/gpu_predict.py [B] [rounds]
Per selected SQP round (QP phase): default launch; the same launch in ORACLE order (descending iteration counts of the default
run: the upper bound of any ordering); stage-1-only launches with a quota (loop states read back: the features); two-stage
launches in index order and in the order the device ranks.  Everything is written to gpurun_out/r2p_predict.npz."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqpsolver_jl_b200.nlp.networks import synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.host.sqp_trust_region import BatchSqpTR, Parameters
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 40
SEL = {7, 12, 20, 30, 38}
QS = tuple(int(q) for q in os.environ.get('QS', '18,22,26,30').split(','))
net = synth_net(118, 186, 54, 118)
pd, qd = net.perturbed_loads(B)
sqp = BatchSqpTR(AcopfPolar(net, pd=pd, qd=qd), B, Parameters(max_iter=rounds, init_mu=1e5), device_evaluator=True)
eng = sqp.optimizer.engine
orig = sqp.optimizer._solve
rec = {}
count = [0]
def hook(phase, x_k, delta, E_override=None, active=None):
    if phase != 0:
        return orig(phase, x_k, delta, E_override, active)
    count[0] += 1
    r = count[0]
    if r not in SEL:
        return orig(phase, x_k, delta, E_override, active)
    def run():
        orig(phase, x_k, delta, E_override, active)
        return sqp.optimizer.last_info.copy(), eng.last_solve_ms
    run()  # warm
    info0, ms0 = run()
    it0 = info0['ipm_iters'].astype(np.int64)
    sel = np.ones(B, bool) if active is None else np.asarray(active, bool)
    rec[f"r{r}_iters"] = it0; rec[f"r{r}_facts"] = info0['chol_factorizations'].astype(np.int64); rec[f"r{r}_active"] = sel
    rec[f"r{r}_status"] = info0['moi_status'].astype(np.int64)
    eng.set_launch_order(np.argsort(-(it0 * sel), kind='stable'))
    info1, ms1 = run()
    eng.set_launch_order(np.argsort(it0 * sel, kind='stable'))
    _, ms1r = run()
    eng.set_launch_order(None)
    same = bool((info1['ipm_iters'] == info0['ipm_iters']).all())
    line = f"round {r:3d} active {int(sel.sum()):4d} iters mean {it0[sel].mean():5.1f} max {it0[sel].max():3d} | default {ms0:6.1f} ms  oracle order {ms1:6.1f}  shortest-first {ms1r:6.1f} (same iters {same})"
    rec[f"r{r}_ms"] = np.array([ms0, ms1, ms1r])
    eng.set_options(method=2)
    for q in QS:
        eng.set_layout(handoff=q, handoff_mode=2)
        _, msq = run()
        st, fl = eng.read_ipm_state()
        rec[f"r{r}_q{q}_state"] = st; rec[f"r{r}_q{q}_flag"] = fl; rec[f"r{r}_q{q}_ms"] = np.array([msq])
        eng.set_layout(handoff=q, handoff_mode=3)
        info3, ms3 = run()
        eng.set_layout(handoff=q, handoff_mode=1)
        info4, ms4 = run()
        ok3 = bool((info3['ipm_iters'][sel] == it0[sel]).all()); ok4 = bool((info4['ipm_iters'][sel] == it0[sel]).all())
        st3 = bool((info3['moi_status'][sel] == info0['moi_status'][sel]).all())
        rec[f"r{r}_q{q}_ms2"] = np.array([ms3, ms4])
        line += f" | q{q}: stage1 {msq:5.1f} two-stage {ms3:6.1f} ranked {ms4:6.1f} (iters same {ok3}/{ok4} status same {st3})"
    eng.set_options(method=0)
    eng.set_layout(handoff=-1, handoff_mode=0)
    print(line, flush=True)
    return orig(phase, x_k, delta, E_override, active)
sqp.optimizer._solve = hook
t0 = time.time(); sqp.run(); print(f"wall {time.time() - t0:.1f} s", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed(os.environ.get("OUT", "gpurun_out/r2p_predict.npz"), **rec)
sqp.close()
