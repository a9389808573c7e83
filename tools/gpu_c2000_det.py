"""DEVELOPMENT TOOL: run the same 2000-bus LP-projection solve several times with ONE interior-point iteration and
compare the engine's work arrays between the runs (hunting nondeterminism)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sqpsolver_jl_b200 import capi
from sqpsolver_jl_b200.nlp.networks import synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
which = sys.argv[1] if len(sys.argv) > 1 else 'c2000'
nlp = AcopfPolar(synth_net(2000, 3000, 400, 2000) if which == 'c2000' else synth_net(118, 186, 54, 118))
x0 = np.asarray(nlp.x0, dtype=float).copy(); X = x0[None, :]
dE = np.zeros((1, nlp.nnz_jac_coo)); nlp.eval_jac_g(X, dE)
E = np.zeros((1, nlp.m)); nlp.eval_g(X, E)
df = np.zeros((1, nlp.n)); nlp.eval_grad_f(X, df)
hv = np.zeros((1, nlp.nnz_hess_coo)); nlp.eval_h(X, 1.0, np.zeros((1, nlp.m)), hv)
N = {k: i for i, k in enumerate("Q XL XU D X ZB YB RB XT R P KP MINV DSH XFIX MASK XW YBW HD TMP TMP2 I1".split())}
M = {k: i for i, k in enumerate("RL RU ES ZC YC RC T RW BC YP YCW TMP AX I1 I2 I3 I4".split())}
snaps = []
for rep in range(int(os.environ.get("REPS", "4"))):
    eng = capi.Engine(0)
    eng.setup_nlp(nlp.n, nlp.m, nlp.num_linear_constraints, nlp.j_row, nlp.j_col, nlp.h_row, nlp.h_col, nlp.x_L, nlp.x_U, nlp.g_L, nlp.g_U)
    eng.set_options(method=2, ipm_max_iter=int(os.environ.get("IPMIT", "1")), team=int(os.environ.get("TEAM", "0")), warm_start=0)
    eng.update_nlp(dE, hv, df, E)
    phase = int(os.environ.get("PHASE", "3"))
    out = eng.solve_tr(phase, X, np.full(1, 10.0))
    print("   info", out[-1][0], flush=True)
    import hashlib
    hs = [hashlib.md5(np.concatenate([a.astype(np.float64).ravel() for a in eng.get_csr(w)]).tobytes()).hexdigest()[:8] for w in (0, 1, 2)]
    print("rep", rep, eng.chol_stats(), "csr hashes", hs, flush=True)
    snap = {"L": eng.debug_read(2, count=2_000_000), "yw": eng.debug_read(3), "dinv": eng.debug_read(4), "wJ": eng.debug_read(5)}
    for k, i in N.items(): snap["N_" + k] = eng.debug_read(0, i)
    for k, i in M.items(): snap["M_" + k] = eng.debug_read(1, i)
    snaps.append(snap)
    eng.close()
ref = snaps[0]
for r, s in enumerate(snaps[1:], 1):
    diffs = []
    for k in ref:
        a, b = ref[k], s[k]
        ok = np.array_equal(a, b, equal_nan=True)
        if not ok:
            bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
            diffs.append((k, int(bad.sum()), int(np.argmax(bad)), float(np.nanmax(np.abs(np.where(bad, a - b, 0.0))))))
    print("run", r, "vs 0:", "identical" if not diffs else diffs)
    if diffs:
        a, b = ref["dinv"][:8], s["dinv"][:8]
        print("   K_jj (1/dinv^2) run0", 1 / a ** 2, "\n   K_jj this run    ", 1 / b ** 2, "\n   diff", 1 / b ** 2 - 1 / a ** 2)
        print("   N_DSH equal", np.array_equal(ref["N_DSH"], s["N_DSH"]), "M_RW equal", np.array_equal(ref["M_RW"], s["M_RW"]), "wJ equal", np.array_equal(ref["wJ"], s["wJ"], equal_nan=True), "N_HD equal", np.array_equal(ref["N_HD"], s["N_HD"]))
        print("   N_DSH[:5]", ref["N_DSH"][:5], "min", np.nanmin(ref["N_DSH"][:16800]))
        break
