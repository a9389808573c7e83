"""DEVELOPMENT TOOL: first-light checks of the CUDA engine on a GPU box (run under gpurun)."""
import sys, os, time, pickle
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sqpsolver_jl_b200 import capi
from sqpsolver_jl_b200.nlp.networks import case9, synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.nlp.toy import ToyExample
from oracle.coo import CooMatrix, SymCooMatrix, csc_pattern
from oracle.sqp_tr import SqpTROracle, Parameters
from oracle.subproblem import trust_region_box
from oracle.qp_solver import kkt_residuals

def check_pattern(nlp, eng):
    rng = np.random.default_rng(0)
    dE = rng.standard_normal(nlp.nnz_jac_coo); hv = rng.standard_normal(nlp.nnz_hess_coo)
    eng.update_nlp(dE, hv, np.zeros(nlp.n), np.zeros(nlp.m))
    Jm = CooMatrix(nlp.j_row, nlp.j_col, nlp.m, nlp.n); Jm.fill(dE)
    Hm = SymCooMatrix(nlp.h_row, nlp.h_col, nlp.n); Hm.fill(hv)
    rp, ci, va = eng.get_csr(0)
    print(' J pattern', np.array_equal(rp, Jm.row_ptr), np.array_equal(ci, Jm.col_idx), 'values bit-exact', np.array_equal(va, Jm.data))
    rp, ci, va = eng.get_csr(2)
    print(' H pattern', np.array_equal(rp, Hm.row_ptr), np.array_equal(ci, Hm.col_idx), 'values bit-exact', np.array_equal(va, Hm.data))
    cp, ri, slot = csc_pattern(nlp.j_row, nlp.j_col, nlp.n)
    rp, ci, va = eng.get_csr(1)
    from oracle.coo import ordered_scatter
    vt = ordered_scatter(slot, np.arange(nlp.nnz_jac_coo), dE, ri.shape[0])
    print(' JT==CSC(J)', np.array_equal(rp[:nlp.n+1], cp), np.array_equal(ci[:cp[-1]], ri), 'values', np.array_equal(va[:cp[-1]], vt))

def run_trace(nlp, trace, eng, only=None):
    Jm = CooMatrix(nlp.j_row, nlp.j_col, nlp.m, nlp.n); Hm = SymCooMatrix(nlp.h_row, nlp.h_col, nlp.n)
    for t in trace:
        if only and t['iter'] not in only: continue
        eng.update_nlp(t['dE'], t['h_val'], t['df'], t['E'])
        t0 = time.time()
        p, lam, mxL, mxU, slack, st, info = eng.solve_tr(capi.PHASE_FR if t['fr'] else capi.PHASE_QP, t['x'], t['Delta'])
        dt = time.time() - t0
        i = info[0]
        line = f"it{t['iter']:3d} {'FR' if t['fr'] else 'QP'} D={t['Delta']:.1e} st={capi.MOI_NAMES.get(int(st[0]), st[0])} orc={t['status']} admm={i['admm_iters']} cg={i['cg_iters']} ipm={i['ipm_iters']} nf={i['chol_factorizations']} ptry={i['polish_tries']} pcg={i['polish_cg_iters']} pol={i['polished']} rho={i['rho']:.1e} rbf={i['rho_box_floor']:.1e} rp={i['res_prim']:.1e} rd={i['res_dual']:.1e} ms={eng.last_solve_ms:.2f}"
        if not t['fr'] and st[0] in (4, 10):
            Jm.fill(t['dE']); Hm.fill(t['h_val'])
            P = Hm.to_scipy(); J = Jm.to_scipy()
            lb, ub = trust_region_box(nlp.x_L - t['x'], nlp.x_U - t['x'], t['Delta'])
            kk = kkt_residuals(P, t['df'], J, nlp.g_L - t['E'], nlp.g_U - t['E'], lb, ub, p[0], lam[0], mxL[0] + mxU[0])
            obj = 0.5 * p[0] @ (P @ p[0]) + t['df'] @ p[0]; obj0 = 0.5 * t['p'] @ (P @ t['p']) + t['df'] @ t['p']
            line += f" |dp|={np.abs(p[0]-t['p']).max():.1e} |dl|={np.abs(lam[0]-t['lambda_qp']).max()/max(1,np.abs(t['lambda_qp']).max()):.1e} dobj={obj-obj0:.1e} kkt=({kk['stationarity']:.1e},{kk['primal']:.1e},{kk['complementarity']:.1e})"
        print(line, flush=True)

if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    eng = capi.Engine(0)
    for a in sys.argv[2:]:
        k, v = a.split('='); eng.set_options(**{k: type(getattr(eng.opts, k))(eval(v))})
    if which in ('all', 'toy'):
        nlp = ToyExample(); trace = []
        s = SqpTROracle(nlp, Parameters(max_iter=100), trace=trace).run(); print('toy oracle', s.status, s.x)
        eng.setup_nlp(nlp.n, nlp.m, nlp.num_linear_constraints, nlp.j_row, nlp.j_col, nlp.h_row, nlp.h_col, nlp.x_L, nlp.x_U, nlp.g_L, nlp.g_U)
        check_pattern(nlp, eng)
        run_trace(nlp, trace, eng)
    if which in ('all', 'case9'):
        nlp = AcopfPolar(case9()); trace = []
        t0 = time.time(); s = SqpTROracle(nlp, Parameters(max_iter=100, init_mu=1e4), trace=trace).run(); print('case9 oracle', s.status, s.obj_val, f'{time.time()-t0:.1f}s')
        eng.setup_nlp(nlp.n, nlp.m, nlp.num_linear_constraints, nlp.j_row, nlp.j_col, nlp.h_row, nlp.h_col, nlp.x_L, nlp.x_U, nlp.g_L, nlp.g_U)
        check_pattern(nlp, eng)
        run_trace(nlp, trace, eng)
    print('launches', eng.launch_count)
