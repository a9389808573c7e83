"""ctypes binding of ``include/sqpqp.h`` (the C-ABI of ``csrc/libsqpqp.so``).

This is what the Python host mirror and the tests call; the Julia shim binds the
very same symbols with ``ccall`` (julia/SqpQpB200.jl).  There is no fallback: if
the shared library is missing the import of :func:`lib` raises, and every compute
call needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
PROF_BUILD = os.environ.get("SQPQP_PROF") == "1"  # development build with the in-kernel phase profile
LIB_PATH = os.path.join(CSRC, "libsqpqp_prof.so" if PROF_BUILD else "libsqpqp.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "sqpqp.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-extended-lambda", "-shared", "-Xcompiler", "-fPIC",
]

# MOI.TerminationStatusCode integers (include/sqpqp.h)
MOI_OPTIMAL = 1
MOI_INFEASIBLE = 2
MOI_LOCALLY_SOLVED = 4
MOI_LOCALLY_INFEASIBLE = 5
MOI_ALMOST_LOCALLY_SOLVED = 10
MOI_ITERATION_LIMIT = 11
MOI_NUMERICAL_ERROR = 20
MOI_NAMES = {
    0: "OPTIMIZE_NOT_CALLED", 1: "OPTIMAL", 2: "INFEASIBLE", 4: "LOCALLY_SOLVED", 5: "LOCALLY_INFEASIBLE",
    10: "ALMOST_LOCALLY_SOLVED", 11: "ITERATION_LIMIT", 20: "NUMERICAL_ERROR",
}
PHASE_QP, PHASE_FR, PHASE_SOC, PHASE_LP = 0, 1, 2, 3
PHASE_MIXED = -1  # host-side convention (QpDevice._solve): QP phase and restoration phase of one round in one call (sqpqp_solve_tr_mixed)


class Info(C.Structure):
    _fields_ = [
        ("moi_status", C.c_int32), ("admm_iters", C.c_int32), ("cg_iters", C.c_int32), ("polish_tries", C.c_int32),
        ("polish_cg_iters", C.c_int32), ("polished", C.c_int32), ("rho_updates", C.c_int32), ("checks", C.c_int32),
        ("ipm_iters", C.c_int32), ("chol_factorizations", C.c_int32),
        ("rho", C.c_double), ("rho_box_floor", C.c_double), ("res_prim", C.c_double), ("res_dual", C.c_double),
        ("objective", C.c_double),
    ]


INFO_DTYPE = np.dtype([
    ("moi_status", "<i4"), ("admm_iters", "<i4"), ("cg_iters", "<i4"), ("polish_tries", "<i4"),
    ("polish_cg_iters", "<i4"), ("polished", "<i4"), ("rho_updates", "<i4"), ("checks", "<i4"),
    ("ipm_iters", "<i4"), ("chol_factorizations", "<i4"),
    ("rho", "<f8"), ("rho_box_floor", "<f8"), ("res_prim", "<f8"), ("res_dual", "<f8"), ("objective", "<f8"),
])
assert INFO_DTYPE.itemsize == C.sizeof(Info)


class Options(C.Structure):
    _fields_ = [
        ("rho0", C.c_double), ("sigma", C.c_double), ("alpha", C.c_double),
        ("eps_abs", C.c_double), ("eps_rel", C.c_double), ("eps_inf", C.c_double),
        ("rho_eq_mult", C.c_double), ("rho_min", C.c_double), ("rho_max", C.c_double), ("adapt_tol", C.c_double),
        ("cg_rel0", C.c_double), ("rb_full_mult", C.c_double),
        ("polish_trigger", C.c_double), ("polish_rho", C.c_double), ("polish_tol", C.c_double),
        ("feas_tol", C.c_double), ("dual_tol", C.c_double),
        ("max_iter", C.c_int32), ("check_every", C.c_int32), ("ruiz_iters", C.c_int32), ("cg_max", C.c_int32),
        ("eig_iters", C.c_int32), ("polish_outer", C.c_int32), ("polish_cg_max", C.c_int32),
        ("warm_start", C.c_int32), ("team", C.c_int32), ("threads", C.c_int32),
        ("method", C.c_int32), ("ipm_max_iter", C.c_int32), ("fallback_max_iter", C.c_int32),
        ("ipm_eps", C.c_double), ("ipm_delta0", C.c_double), ("ipm_delta_min", C.c_double), ("ipm_rho0", C.c_double),
        ("ipm_tau", C.c_double), ("ipm_mu0", C.c_double), ("ipm_mu_min", C.c_double), ("ipm_kappa_eps", C.c_double),
        ("ipm_refine", C.c_int32), ("verbose", C.c_int32), ("occupancy", C.c_int32), ("smem_kb", C.c_int32),
        ("ipm_ic_growth", C.c_double), ("ipm_ic_decay", C.c_double),
    ]


EXPORTS = [
    "sqpqp_create", "sqpqp_destroy", "sqpqp_last_error", "sqpqp_default_options", "sqpqp_set_options", "sqpqp_stream",
    "sqpqp_setup_nlp", "sqpqp_update_nlp", "sqpqp_update_nlp_device", "sqpqp_solve_tr", "sqpqp_num_slacks",
    "sqpqp_merit", "sqpqp_kt_residuals", "sqpqp_jac_times", "sqpqp_get_csr", "sqpqp_qp_setup", "sqpqp_qp_solve",
    "sqpqp_launch_count", "sqpqp_last_solve_ms", "sqpqp_last_solve_kernel", "sqpqp_solve_tr_device", "sqpqp_sync", "sqpqp_device_outputs",
    "sqpqp_fetch_info", "sqpqp_chol_stats", "sqpqp_chol_layout", "sqpqp_prof_read", "sqpqp_spmv", "sqpqp_spmv_device", "sqpqp_debug_read", "sqpqp_debug_set", "sqpqp_linesearch_terms", "sqpqp_acopf_setup", "sqpqp_acopf_eval_update",
    "sqpqp_host_register", "sqpqp_host_unregister", "sqpqp_solve_ms_total", "sqpqp_acopf_eval_trial",
    "sqpqp_set_launch_order", "sqpqp_debug_read_state", "sqpqp_solve_tr_mixed", "sqpqp_solve_tr_mixed_device",
]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/sqpqp.cu for sm_100a with nvcc (cross-compiles without a GPU).
    SQPQP_PROF=1 in the environment adds the in-kernel phase profile (development builds only)."""
    import fcntl

    def stale():
        srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp", ".h"))] + [HEADER]
        return not (os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs))

    if not force and not stale():
        return LIB_PATH
    # several ranks of one node may get here at once (torchrun): one compiles, the others wait on the lock and find the
    # library fresh; the compiler writes a temporary file that is renamed into place, so a reader never maps a partial .so
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or stale():
                nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
                tmp = LIB_PATH + ".tmp.%d" % os.getpid()
                cmd = ([nvcc] + NVCC_FLAGS + (["-DSQPQP_PROF"] if PROF_BUILD else []) + (["-Xptxas", "-v"] if verbose else [])
                       + ["-o", tmp, os.path.join(CSRC, "sqpqp.cu")])
                res = subprocess.run(cmd, capture_output=True, text=True)
                if res.returncode != 0:
                    if os.path.exists(tmp):
                        os.remove(tmp)
                    raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
                os.replace(tmp, LIB_PATH)
                if verbose:
                    sys.stderr.write(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


def lib():
    """Load libsqpqp.so (must have been built: ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the engine has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.sqpqp_create.argtypes = [C.POINTER(vp), C.c_int]
    L.sqpqp_destroy.argtypes = [vp]
    L.sqpqp_last_error.argtypes = [vp]
    L.sqpqp_last_error.restype = C.c_char_p
    L.sqpqp_default_options.argtypes = [C.POINTER(Options)]
    L.sqpqp_default_options.restype = None
    L.sqpqp_set_options.argtypes = [vp, C.POINTER(Options)]
    L.sqpqp_stream.argtypes = [vp]
    L.sqpqp_stream.restype = vp
    L.sqpqp_setup_nlp.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, _lp, _lp, C.c_int64, _lp, _lp,
                                  _dp, _dp, _dp, _dp, C.c_int32]
    L.sqpqp_update_nlp.argtypes = [vp, _dp, _dp, _dp, _dp]
    L.sqpqp_update_nlp_device.argtypes = [vp, vp, vp, vp, vp]
    L.sqpqp_solve_tr.argtypes = [vp, C.c_int32, _dp, _dp, _dp, _ip, _dp, _dp, _dp, _dp, _dp, _ip, C.c_void_p]
    L.sqpqp_solve_tr_device.argtypes = [vp, C.c_int32, vp, vp, vp, vp]
    L.sqpqp_solve_tr_mixed.argtypes = [vp, _dp, _dp, _ip, _ip, _dp, _dp, _dp, _dp, _dp, _ip, C.c_void_p]
    L.sqpqp_solve_tr_mixed_device.argtypes = [vp, vp, vp, vp, vp]
    L.sqpqp_sync.argtypes = [vp]
    L.sqpqp_device_outputs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.sqpqp_fetch_info.argtypes = [vp, C.c_void_p]
    L.sqpqp_chol_stats.argtypes = [vp, _lp, _lp, _lp]
    L.sqpqp_chol_layout.argtypes = [vp, _lp, _lp]
    L.sqpqp_prof_read.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.sqpqp_acopf_setup.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _ip, _ip, _ip, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp,
                                    C.c_int32, _ip, _ip, _ip, _dp, C.c_int32, _ip]
    L.sqpqp_acopf_eval_update.argtypes = [vp, _dp, _dp, _ip, _dp, _dp, _dp]
    L.sqpqp_linesearch_terms.argtypes = [vp, _dp, _dp, _dp, _dp, _dp, _dp, _dp]
    L.sqpqp_debug_set.argtypes = [vp, C.c_int32, C.c_int32]
    L.sqpqp_acopf_eval_trial.argtypes = [vp, _dp, _ip, _dp, _dp]
    L.sqpqp_host_register.argtypes = [vp, vp, C.c_int64]
    L.sqpqp_host_unregister.argtypes = [vp, vp]
    L.sqpqp_debug_read.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, _dp, C.c_int64]
    L.sqpqp_set_launch_order.argtypes = [vp, _ip]
    L.sqpqp_debug_read_state.argtypes = [vp, vp, C.c_int64]
    L.sqpqp_spmv.argtypes = [vp, C.c_int32, _dp, _dp]
    L.sqpqp_spmv_device.argtypes = [vp, C.c_int32, vp, vp]
    L.sqpqp_num_slacks.argtypes = [vp, _ip]
    L.sqpqp_merit.argtypes = [vp, _dp, _dp, _dp, _dp, _dp, _ip, _dp, _dp, _dp, _dp, _dp]
    L.sqpqp_kt_residuals.argtypes = [vp, _dp, _dp, _dp, _dp]
    L.sqpqp_jac_times.argtypes = [vp, _dp, _dp]
    L.sqpqp_get_csr.argtypes = [vp, C.c_int32, C.c_int32, _lp, _ip, _ip, _dp]
    L.sqpqp_qp_setup.argtypes = [vp, C.c_int32, C.c_int32, C.c_int64, _lp, _lp, C.c_int64, _lp, _lp]
    L.sqpqp_qp_solve.argtypes = [vp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _dp, _ip, C.c_void_p]
    L.sqpqp_launch_count.argtypes = [vp]
    L.sqpqp_launch_count.restype = C.c_int64
    L.sqpqp_last_solve_ms.argtypes = [vp]
    L.sqpqp_last_solve_ms.restype = C.c_double
    L.sqpqp_solve_ms_total.argtypes = [vp]
    L.sqpqp_solve_ms_total.restype = C.c_double
    L.sqpqp_last_solve_kernel.argtypes = [vp]
    L.sqpqp_last_solve_kernel.restype = C.c_char_p
    for name in EXPORTS:
        f = getattr(L, name)
        if f.restype is C.c_int:  # default restype
            f.restype = C.c_int
    _lib = L
    return L


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _l(a):
    return None if a is None else a.ctypes.data_as(_lp)


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != shape:
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


class SqpQpError(RuntimeError):
    pass


class Engine:
    """Thin object wrapper over one ``sqpqp_handle`` (one GPU, one stream)."""

    def __init__(self, device: int = 0):
        self.L = lib()
        self.h = C.c_void_p()
        rc = self.L.sqpqp_create(C.byref(self.h), device)
        if rc != 0:
            raise SqpQpError(f"sqpqp_create failed with {rc} (no CUDA device {device}? the engine has no CPU fallback)")
        self.opts = Options()
        self.L.sqpqp_default_options(C.byref(self.opts))
        self.batch = self.n = self.m = self.S = 0
        # True: solve_tr writes into one persistent set of result arrays (valid until the next call) instead of
        # allocating new ones -- for a batched host that owns its result buffers (bench.py's e2e leg)
        self.reuse_outputs = False
        self._out = None
        self._registered = []  # keeps registered arrays alive
        self.register_outputs = False  # with reuse_outputs: page-lock the persistent result arrays too

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.L.sqpqp_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise SqpQpError(f"sqpqp error {rc}: {self.L.sqpqp_last_error(self.h).decode()}")

    def register_host(self, *arrays):
        """Page-lock caller-owned numpy arrays (sqpqp_host_register): later calls that are handed these arrays copy straight
        between them and the device.  The arrays must stay alive (and must not be reallocated) until close()."""
        for a in arrays:
            if a is None or a.nbytes == 0:
                continue
            assert a.flags["C_CONTIGUOUS"]
            self._ck(self.L.sqpqp_host_register(self.h, C.c_void_p(a.ctypes.data), a.nbytes))
            self._registered.append(a)

    def set_options(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.opts, k):
                raise KeyError(k)
            setattr(self.opts, k, v)
        self._ck(self.L.sqpqp_set_options(self.h, C.byref(self.opts)))

    def set_launch_order(self, order=None):
        """CTA slot k of the following batched solves runs instance order[k] (None: index order)."""
        if order is None:
            self._ck(self.L.sqpqp_set_launch_order(self.h, None))
        else:
            o = np.ascontiguousarray(order, dtype=np.int32)
            assert o.shape == (self.batch,)
            self._ck(self.L.sqpqp_set_launch_order(self.h, _i(o)))

    IPM_STATE_DTYPE = np.dtype([("delta", "f8"), ("rho_p", "f8"), ("rho_last", "f8"), ("mu_t", "f8"), ("alpha", "f8"), ("sig_prev", "f8"),
                                ("del_prev", "f8"), ("rp_ref", "f8"), ("nin", "f8"), ("it", "i4"), ("nfact", "i4"), ("acc_cnt", "i4"), ("pad", "i4")])

    def read_ipm_state(self):
        """(development) loop states saved by a launch with an iteration quota, and the per-instance flags (3 = stopped by the quota)."""
        B = self.batch
        raw = np.zeros(B * (self.IPM_STATE_DTYPE.itemsize + 4), dtype=np.uint8)
        self._ck(self.L.sqpqp_debug_read_state(self.h, raw.ctypes.data_as(C.c_void_p), raw.nbytes))
        st = raw[:B * self.IPM_STATE_DTYPE.itemsize].view(self.IPM_STATE_DTYPE).copy()
        fl = raw[B * self.IPM_STATE_DTYPE.itemsize:].view(np.int32).copy()
        return st, fl

    def set_layout(self, G=0, threads=0, ctas_per_sm=0, tail=-1, ring=None, fuse=None, handoff=None, handoff_mode=None):
        """Layout of the batched interior-point launch, effective at the next setup_nlp: G instances interleaved per CTA
        (0 = auto, 1 = one CTA per instance, 2 / 4 / 8), its CTA size (0 = auto, 256 / 512 / 1024), CTAs per SM
        (0 = auto) and the cap of the dense tail of the factor in columns (-1 = auto).  ring (effective at the next solve):
        0 = auto, 1 = never stream the index programs through the shared-memory ring, 2 = whenever they exist.
        Tuning / A-B runs."""
        for what, v in ((2, G), (3, threads), (4, ctas_per_sm), (1, tail)):
            self._ck(self.L.sqpqp_debug_set(self.h, what, int(v)))
        if ring is not None:
            self._ck(self.L.sqpqp_debug_set(self.h, 6, int(ring)))
        if handoff is not None:  # iteration quota before the resident launch takes an instance over (-1 auto, 0 off)
            self._ck(self.L.sqpqp_debug_set(self.h, 8, int(handoff)))
        if handoff_mode is not None:  # 0 resident launch takes the stragglers, 1 second throughput launch longest-predicted-first, 2 stop, 3 second launch in index order
            self._ck(self.L.sqpqp_debug_set(self.h, 9, int(handoff_mode)))
        if fuse is not None:  # (next setup_nlp) 0: forward sweep of the Newton solve as its own level-scheduled pass
            self._ck(self.L.sqpqp_debug_set(self.h, 7, int(fuse)))

    # ---- NLP lane --------------------------------------------------------------
    def setup_nlp(self, n, m, m_lin, j_row, j_col, h_row, h_col, x_L, x_U, g_L, g_U, batch=1):
        j_row = np.ascontiguousarray(j_row, dtype=np.int64)
        j_col = np.ascontiguousarray(j_col, dtype=np.int64)
        h_row = np.ascontiguousarray(h_row if h_row is not None else [], dtype=np.int64)
        h_col = np.ascontiguousarray(h_col if h_col is not None else [], dtype=np.int64)
        x_L, x_U, g_L, g_U = (_f64(a) for a in (x_L, x_U, g_L, g_U))
        per = int(g_L.ndim == 2 or x_L.ndim == 2)
        if per:
            x_L = np.ascontiguousarray(np.broadcast_to(x_L, (batch, n)))
            x_U = np.ascontiguousarray(np.broadcast_to(x_U, (batch, n)))
            g_L = np.ascontiguousarray(np.broadcast_to(g_L, (batch, m)))
            g_U = np.ascontiguousarray(np.broadcast_to(g_U, (batch, m)))
        self._ck(self.L.sqpqp_setup_nlp(self.h, batch, n, m, m_lin, j_row.shape[0], _l(j_row), _l(j_col), h_row.shape[0],
                                        _l(h_row), _l(h_col), _d(x_L), _d(x_U), _d(g_L), _d(g_U), per))
        self.batch, self.n, self.m = batch, n, m
        self.nnz_j, self.nnz_h = j_row.shape[0], h_row.shape[0]
        s = C.c_int32()
        self._ck(self.L.sqpqp_num_slacks(self.h, C.byref(s)))
        self.S = s.value

    def update_nlp(self, dE, h_val, df, E):
        B = self.batch
        dE = _f64(dE).reshape(B, self.nnz_j)
        h_val = _f64(h_val).reshape(B, self.nnz_h) if self.nnz_h else None
        df = _f64(df).reshape(B, self.n)
        E = _f64(E).reshape(B, self.m)
        self._ck(self.L.sqpqp_update_nlp(self.h, _d(dE), _d(h_val), _d(df), _d(E)))

    def update_nlp_device(self, dE_ptr, h_val_ptr, df_ptr, E_ptr):
        self._ck(self.L.sqpqp_update_nlp_device(self.h, dE_ptr, h_val_ptr, df_ptr, E_ptr))

    def solve_tr(self, phase, x_k, delta, E_override=None, active=None, mixed=None):
        B, n, m, S = self.batch, self.n, self.m, self.S
        x_k = _f64(x_k).reshape(B, n)
        delta = np.ascontiguousarray(np.broadcast_to(np.asarray(delta, dtype=np.float64), (B,)))
        E_override = _f64(E_override).reshape(B, m) if E_override is not None else None
        act = np.ascontiguousarray(active, dtype=np.int32).reshape(B) if active is not None else None
        if (self.reuse_outputs and self._out is not None and self._out[0].shape == (B, n) and self._out[1].shape == (B, m)
                and self._out[4].shape == (B, max(S, 1))):
            # caller-owned result buffers, as a compiled host would pass to the C ABI: the arrays returned by the
            # previous call are overwritten (fresh 60 MB numpy arrays cost ~6 ms of page faults per call at batch 1024)
            p, lam, mxL, mxU, slack, status, info = self._out
        else:
            p = np.zeros((B, n))
            lam = np.zeros((B, m))
            mxL = np.zeros((B, n))
            mxU = np.zeros((B, n))
            slack = np.zeros((B, max(S, 1)))
            status = np.zeros(B, dtype=np.int32)
            info = np.zeros(B, dtype=INFO_DTYPE)
            if self.reuse_outputs:
                self._out = (p, lam, mxL, mxU, slack, status, info)
                if self.register_outputs:
                    self.register_host(p, lam, mxL, mxU, slack, status, info)
        if mixed is not None:
            aq, af = (np.ascontiguousarray(a, dtype=np.int32).reshape(B) for a in mixed)
            self._ck(self.L.sqpqp_solve_tr_mixed(self.h, _d(x_k), _d(delta), _i(aq), _i(af), _d(p), _d(lam), _d(mxL), _d(mxU),
                                                 _d(slack), _i(status), info.ctypes.data_as(C.c_void_p)))
        else:
            self._ck(self.L.sqpqp_solve_tr(self.h, phase, _d(x_k), _d(delta), _d(E_override), _i(act), _d(p), _d(lam), _d(mxL),
                                           _d(mxU), _d(slack), _i(status), info.ctypes.data_as(C.c_void_p)))
        return p, lam, mxL, mxU, slack[:, :S], status, info

    def solve_tr_mixed(self, x_k, delta, active_qp, active_fr):
        """One SQP round of a batch with instances in both phases (sqpqp_solve_tr_mixed): QP subproblems over `active_qp`,
        restoration LPs over `active_fr` (disjoint), launched side by side; one set of results."""
        return self.solve_tr(None, x_k, delta, mixed=(active_qp, active_fr))

    def solve_tr_mixed_device(self, x_k_ptr, delta_ptr, active_qp_ptr, active_fr_ptr):
        self._ck(self.L.sqpqp_solve_tr_mixed_device(self.h, x_k_ptr, delta_ptr, active_qp_ptr, active_fr_ptr))

    def solve_tr_device(self, phase, x_k_ptr, delta_ptr, E_override_ptr=None, active_ptr=None):
        """Device-pointer, non-blocking variant (ints are raw device addresses)."""
        self._ck(self.L.sqpqp_solve_tr_device(self.h, phase, x_k_ptr, delta_ptr, E_override_ptr, active_ptr))

    def sync(self):
        self._ck(self.L.sqpqp_sync(self.h))

    def fetch_info(self):
        info = np.zeros(self.batch, dtype=INFO_DTYPE)
        self._ck(self.L.sqpqp_fetch_info(self.h, info.ctypes.data_as(C.c_void_p)))
        return info

    def chol_stats(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self.L.sqpqp_chol_stats(self.h, C.byref(a), C.byref(b), C.byref(c)))
        t, lv = C.c_int64(), C.c_int64()
        self._ck(self.L.sqpqp_chol_layout(self.h, C.byref(t), C.byref(lv)))
        return {"nnzL": a.value, "levels": b.value, "flops": c.value, "dense_tail": t.value, "tree_levels": lv.value}

    def spmv(self, which, x):
        """y = J x (which=0), J' x (1), H x (2) per instance on the device matrices; host arrays in and out."""
        B = self.batch
        nx = self.m if which == 1 else self.n
        ny = self.m if which == 0 else self.n
        x = _f64(x).reshape(B, nx)
        y = np.empty((B, ny))
        self._ck(self.L.sqpqp_spmv(self.h, which, _d(x), _d(y)))
        return y

    def spmv_device(self, which, x_dev_ptr, y_dev_ptr):
        self._ck(self.L.sqpqp_spmv_device(self.h, which, C.c_void_p(int(x_dev_ptr)), C.c_void_p(int(y_dev_ptr))))

    def debug_read(self, kind, idx=0, b=0, count=None):
        count = int(count if count is not None else max(self.n + self.S, self.m, 1) * 64)
        out = np.full(count, np.nan)
        self._ck(self.L.sqpqp_debug_read(self.h, kind, idx, b, _d(out), count))
        return out

    def prof_read(self):
        """Cycles per solve segment (csrc/common.cuh ProfSeg) since the last call; zeros in a normal build."""
        out = (C.c_uint64 * 32)()
        self._ck(self.L.sqpqp_prof_read(self.h, out))
        return np.array(list(out), dtype=np.uint64)

    @property
    def stream(self):
        return int(self.L.sqpqp_stream(self.h) or 0)

    def merit(self, x, p, E_trial, f_trial, mu, fr=None):
        B, n, m = self.batch, self.n, self.m
        x = _f64(x).reshape(B, n)
        p = _f64(p).reshape(B, n)
        # None: the trial values acopf_eval_trial left on the device
        E_trial = _f64(E_trial).reshape(B, m) if E_trial is not None else None
        f_trial = np.ascontiguousarray(np.broadcast_to(np.asarray(f_trial, dtype=np.float64), (B,))) if f_trial is not None else None
        mu = np.ascontiguousarray(np.broadcast_to(np.asarray(mu, dtype=np.float64), (B,)))
        frv = np.ascontiguousarray(np.broadcast_to(np.asarray(fr, dtype=np.int32), (B,))) if fr is not None else None
        out = [np.zeros(B) for _ in range(5)]
        self._ck(self.L.sqpqp_merit(self.h, _d(x), _d(p), _d(E_trial), _d(f_trial), _d(mu), _i(frv), *[_d(o) for o in out]))
        return dict(viol0=out[0], viol_trial=out[1], phi_trial=out[2], q0=out[3], qk=out[4])

    def kt_residuals(self, lam, mult_x_U, mult_x_L):
        B, n, m = self.batch, self.n, self.m
        lam = _f64(lam).reshape(B, m)
        mxU = _f64(mult_x_U).reshape(B, n)
        mxL = _f64(mult_x_L).reshape(B, n)
        kt = np.zeros(B)
        self._ck(self.L.sqpqp_kt_residuals(self.h, _d(lam), _d(mxU), _d(mxL), _d(kt)))
        return kt

    def acopf_setup(self, nlp):
        """Hand the network of an :class:`AcopfPolar` NLP (the one given to setup_nlp) to the device-side evaluator."""
        net = nlp.net
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        nb, ng, nl = nlp.nb, nlp.ng, nlp.nl
        # balance rows as CSR over their Jacobian COO entries (acopf.py: _build_jacobian_structure)
        ptr, col, kind = [0], [], []
        for i in range(nb):
            for part, (ov, og) in enumerate(((nlp.o_p, nlp.o_pg), (nlp.o_q, nlp.o_qg))):
                for a in nlp.arcs_at[i]:
                    col.append(ov + a); kind.append(0)
                for k in nlp.gens_at[i]:
                    col.append(og + k); kind.append(0)
                if nlp.has_shunt[i]:
                    col.append(nlp.o_vm + i); kind.append(1 if part == 0 else 2)
                ptr.append(len(col))
        keep = (f64(net.f_bus), )  # noqa: F841  (arrays below are copied by the library during the call)
        args = [i32(net.f_bus), i32(net.t_bus), i32(net.gen_bus), f64(nlp.oa), f64(nlp.oc), f64(nlp.os), f64(net.cost2),
                f64(net.cost1), f64(net.cost0), f64(net.gs), f64(net.bs)]
        bal = [i32(ptr), i32(col), i32(kind), f64(nlp.j_bal_const), i32(nlp.sh_bus)]
        assert len(col) == nlp.j_bal_const.shape[0]
        self._ck(self.L.sqpqp_acopf_setup(self.h, nb, ng, nl, int(net.ref_bus), _i(args[0]), _i(args[1]), _i(args[2]), _d(args[3]),
                                          _d(args[4]), _d(args[5]), _d(args[6]), _d(args[7]), _d(args[8]), _d(args[9]), _d(args[10]),
                                          len(col), _i(bal[0]), _i(bal[1]), _i(bal[2]), _d(bal[3]), int(nlp.sh_bus.shape[0]),
                                          _i(bal[4])))

    def acopf_eval_update(self, x, lam, mask=None):
        """eval_functions! on the device for the masked instances + scatter; returns (f[B], E[B,m], df[B,n])."""
        B, n, m = self.batch, self.n, self.m
        x = _f64(x).reshape(B, n); lam = _f64(lam).reshape(B, m)
        mk = np.ascontiguousarray(mask, dtype=np.int32) if mask is not None else None
        f = np.zeros(B); E = np.zeros((B, m)); df = np.zeros((B, n))
        self._ck(self.L.sqpqp_acopf_eval_update(self.h, _d(x), _d(lam), _i(mk) if mk is not None else None, _d(f), _d(E), _d(df)))
        return f, E, df

    def acopf_eval_trial(self, x_trial, mask=None, fetch=False):
        """f and g at a trial point on the device (function values only), kept there for merit(..., None, None, ...);
        fetch=True also returns (f[B], E[B,m])."""
        B, n, m = self.batch, self.n, self.m
        x = _f64(x_trial).reshape(B, n)
        mk = np.ascontiguousarray(mask, dtype=np.int32) if mask is not None else None
        f = np.zeros(B) if fetch else None
        E = np.zeros((B, m)) if fetch else None
        self._ck(self.L.sqpqp_acopf_eval_trial(self.h, _d(x), _i(mk) if mk is not None else None, _d(f), _d(E)))
        return (f, E) if fetch else None

    def linesearch_terms(self, x, p, alpha, E_trial, mu_rows, lam):
        """Device line-search primitives (include/sqpqp.h: sqpqp_linesearch_terms); returns a dict of [batch] arrays."""
        B, n, m = self.batch, self.n, self.m
        x = _f64(x).reshape(B, n); p = _f64(p).reshape(B, n)
        alpha = np.ascontiguousarray(np.broadcast_to(np.asarray(alpha, dtype=np.float64), (B,)))
        E_trial = _f64(E_trial).reshape(B, m); mu_rows = _f64(mu_rows).reshape(B, m); lam = _f64(lam).reshape(B, m)
        out = np.zeros((8, B))
        self._ck(self.L.sqpqp_linesearch_terms(self.h, _d(x), _d(p), _d(alpha), _d(E_trial), _d(mu_rows), _d(lam), _d(out)))
        keys = ("dfp", "pHp", "viol1", "violinf", "wviol0", "wviol_trial", "viol1_trial", "compl")
        return {k: out[i] for i, k in enumerate(keys)}

    def jac_times(self, p):
        p = _f64(p).reshape(self.batch, self.n)
        out = np.zeros((self.batch, self.m))
        self._ck(self.L.sqpqp_jac_times(self.h, _d(p), _d(out)))
        return out

    def get_csr(self, which, b=0):
        nnz = C.c_int64()
        self._ck(self.L.sqpqp_get_csr(self.h, which, b, C.byref(nnz), None, None, None))
        nrows = self.m if which == 0 else self.n
        rp = np.zeros(nrows + 1, dtype=np.int32)
        ci = np.zeros(max(nnz.value, 1), dtype=np.int32)
        va = np.zeros(max(nnz.value, 1))
        self._ck(self.L.sqpqp_get_csr(self.h, which, b, C.byref(nnz), _i(rp), _i(ci), _d(va)))
        return rp, ci[: nnz.value], va[: nnz.value]

    # ---- generic lane ----------------------------------------------------------
    def qp_setup(self, nv, nc, p_row, p_col, a_row, a_col):
        p_row = np.ascontiguousarray(p_row, dtype=np.int64)
        p_col = np.ascontiguousarray(p_col, dtype=np.int64)
        a_row = np.ascontiguousarray(a_row, dtype=np.int64)
        a_col = np.ascontiguousarray(a_col, dtype=np.int64)
        self._ck(self.L.sqpqp_qp_setup(self.h, nv, nc, p_row.shape[0], _l(p_row), _l(p_col), a_row.shape[0], _l(a_row), _l(a_col)))
        self.batch, self.n, self.m, self.S = 1, nv, nc, 0
        self.nnz_j, self.nnz_h = a_row.shape[0], p_row.shape[0]

    def qp_solve(self, p_val, q, a_val, rl, ru, cl, cu):
        x = np.zeros(self.n)
        rd = np.zeros(max(self.m, 1))
        cd = np.zeros(self.n)
        st = C.c_int32()
        info = np.zeros(1, dtype=INFO_DTYPE)
        self._ck(self.L.sqpqp_qp_solve(self.h, _d(_f64(p_val)), _d(_f64(q)), _d(_f64(a_val)), _d(_f64(rl)), _d(_f64(ru)),
                                       _d(_f64(cl)), _d(_f64(cu)), _d(x), _d(rd), _d(cd), C.byref(st),
                                       info.ctypes.data_as(C.c_void_p)))
        return x, rd[: self.m], cd, st.value, info[0]

    @property
    def launch_count(self):
        return int(self.L.sqpqp_launch_count(self.h))

    @property
    def last_solve_kernel(self):
        return self.L.sqpqp_last_solve_kernel(self.h).decode()

    @property
    def last_solve_ms(self):
        return float(self.L.sqpqp_last_solve_ms(self.h))

    @property
    def solve_ms_total(self):
        """CUDA-event time of all solve launches of this handle so far (ms)."""
        return float(self.L.sqpqp_solve_ms_total(self.h))
