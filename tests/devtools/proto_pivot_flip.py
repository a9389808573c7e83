"""DEVELOPMENT TOOL ONLY (CPU) -- local inertia correction in the interior point: non-positive pivots flipped in place
(d -> |d|, a diagonal correction on the offending pivots only) instead of a larger scalar shift and a new factorisation.
Runs tests/devtools/proto_ipm.py with a dense LDL' in the DEVICE's elimination order (proto_pivot_flip.cpp) on subproblems
recorded on the GPU by tools/gpu_record_slow.py (npz path = argv[3], default gpurun_out/r2r_slow_qps.npz).
usage: python proto_pivot_flip.py VARIANT [i,j,...] [npz]     VARIANT: base | abs<floor> | hyb<M>[_<floor>[_<growth>[_<decay>]]]
Outcome (profiles/r02_tuning.md section 9): which instances are slow is a property of the TRAJECTORY, not of the subproblem --
on subproblems recorded as the slowest of one policy every other policy looks 40 % better; measured on the whole batch on the
device the policies are equal in the mean and 'hyb1' has the worse tail.  Not built into the kernels."""
import sys, time, ctypes as C
import numpy as np, scipy.sparse as sp
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/tests/devtools')
import proto_ipm as PI
from support.closed_loop import qp_of_trace, scaled_kkt
from sqpsolver_jl_b200.nlp.networks import synth_net
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
import os, subprocess
_here = os.path.dirname(os.path.abspath(__file__))
_so = '/tmp/libproto_pivot_flip.so'
if not os.path.exists(_so) or os.path.getmtime(_so) < os.path.getmtime(os.path.join(_here, 'proto_pivot_flip.cpp')):
    subprocess.check_call(['g++', '-O3', '-march=native', '-std=c++17', '-shared', '-fPIC', '-o', _so, os.path.join(_here, 'proto_pivot_flip.cpp')])
lib = C.CDLL(_so)
dp = C.POINTER(C.c_double); ip = C.POINTER(C.c_int32)
d = np.load(sys.argv[3] if len(sys.argv) > 3 else '/root/repo/gpurun_out/r2r_slow_qps.npz')
ROUNDS = sorted({int(k[1:].split('_')[0]) for k in d.files if k.endswith('_ids')})
net = synth_net(118, 186, 54, 118)
nlp0 = AcopfPolar(net)
PERM = [None]
class ModLU:
    def __init__(self, K, mode, thresh, flo):
        n = K.shape[0]
        if PERM[0] is None:
            PERM[0] = np.arange(n)
        p = PERM[0]
        A = np.ascontiguousarray(K.toarray()[np.ix_(p, p)])
        nmod = C.c_int(0); mp = C.c_double(0)
        rc = lib.ldl_mod(A.ctypes.data_as(dp), n, mode, C.c_double(thresh), C.c_double(flo), C.byref(nmod), C.byref(mp))
        self.ok = rc == 0; self.A = A; self.n = n; self.nmod = nmod.value; self.p = p; self.minpiv = mp.value
    def solve(self, rhs):
        x = np.ascontiguousarray(rhs[self.p], dtype=np.float64).copy()
        lib.ldl_solve(self.A.ctypes.data_as(dp), self.n, x.ctypes.data_as(dp))
        out = np.empty_like(x); out[self.p] = x
        return out
def set_perm(P, J):
    P = sp.csr_matrix(P); J = sp.csr_matrix(J); P.sort_indices(); J.sort_indices()
    n = P.shape[0]; perm = np.zeros(n, np.int32)
    a = lambda v: np.ascontiguousarray(v, dtype=np.int32)
    jrp, jc, prp, pc = a(J.indptr), a(J.indices), a(P.indptr), a(P.indices)
    rc = lib.sym_perm(n, J.shape[0], jrp.ctypes.data_as(ip), jc.ctypes.data_as(ip), prp.ctypes.data_as(ip), pc.ctypes.data_as(ip), 96, perm.ctypes.data_as(ip))
    assert rc == 0
    PERM[0] = perm.astype(np.int64)
def qps():
    for r in ROUNDS:
        ids = d[f'r{r}_ids']
        for k, b in enumerate(ids):
            nlp = AcopfPolar(net, pd=d['pd'][b], qd=d['qd'][b])
            t = {n_: d[f'r{r}_{n_}'][k] for n_ in ('x', 'Delta', 'dE', 'h_val', 'df', 'E')}
            yield r, int(b), int(d[f'r{r}_iters'][k]), int(d[f'r{r}_facts'][k]), float(d[f'r{r}_obj'][k]), qp_of_trace(nlp, t)

def run(variant, items, verbose=False):
    tot_it = tot_f = 0
    for r, b, dev_it, dev_f, dev_obj, (P, q, J, rl, ru, xl, xu) in items:
        if PERM[0] is None or len(PERM[0]) != P.shape[0]:
            set_perm(P, J)
        o = PI.Opts(); o.max_iter = 200; o.delta_min = 1e-8; o.rho_bump = 4.0; o.rho_dec = 3.0; o.refine = False
        nm = [0]
        if variant == 'base':
            o.factor = lambda K: (lambda f: (f, f.ok))(ModLU(K, 0, 0.0, 0.0))
        elif variant.startswith('abs'):   # |d| with floor
            flo = float(variant[3:] or 1e-8)
            def fac(K, flo=flo):
                f = ModLU(K, 1, 0.0, flo); nm[0] += f.nmod; return f, True
            o.factor = fac
        elif variant.startswith('hyb'):   # |d| accepted while at most M pivots were flipped, else the shift is raised
            parts = variant[3:].split('_')
            M = int(parts[0]); flo = float(parts[1]) if len(parts) > 1 else 1e-8
            if len(parts) > 2: o.rho_bump = float(parts[2])
            if len(parts) > 3: o.rho_dec = float(parts[3])
            def fac(K, M=M, flo=flo):
                f = ModLU(K, 1, 0.0, flo); nm[0] += f.nmod if f.nmod <= M else 0; return f, f.nmod <= M
            o.factor = fac
        elif variant.startswith('fix'):
            flo = float(variant[3:])
            def fac(K, flo=flo):
                f = ModLU(K, 2, 0.0, flo); nm[0] += f.nmod; return f, True
            o.factor = fac
        t0 = time.time()
        res = PI.ipm_solve(P, q, J, rl, ru, xl, xu, o)
        x = res['x']; obj = 0.5 * x @ (P @ x) + q @ x
        kkt = scaled_kkt(P, q, J, rl, ru, xl, xu, x, res['yc'], res['yb']) if 'yc' in res else np.nan
        it, nf = res['info']['iters'], res['info']['nfact']
        tot_it += it; tot_f += nf
        print(f"  {variant:8s} r{r} b{b:4d} device {dev_it:3d}/{dev_f:3d} obj {dev_obj:+.6e} | proto {res['status']:8s} it {it:3d} fact {nf:3d} mods {nm[0]:4d} obj {obj:+.6e} kkt {kkt:.1e}  ({time.time()-t0:.0f}s)", flush=True)
    print(f"{variant}: total iterations {tot_it} factorisations {tot_f}", flush=True)

if __name__ == "__main__":
    items = list(qps())
    sel = [int(a) for a in sys.argv[2].split(',')] if len(sys.argv) > 2 else range(len(items))
    run(sys.argv[1], [items[k] for k in sel])
