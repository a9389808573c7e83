"""Host-only check of the sparse-Cholesky symbolic analysis (csrc/symbolic.hpp).

The index programs (assembly, level-scheduled factorisation, triangular sweeps) are executed
on the CPU by a test-only helper (tests/support/symbolic_check.cpp) in the order the device
kernels execute them and compared with a dense solve of  K = P + diag(d) + J' diag(w) J.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from oracle.coo import CooMatrix, SymCooMatrix
from sqpsolver_jl_b200.nlp.acopf import AcopfPolar
from sqpsolver_jl_b200.nlp.networks import case9, synth_net

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def symlib():
    src = os.path.join(HERE, "support", "symbolic_check.cpp")
    out = os.path.join(HERE, "support", "libsymcheck.so")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(
            os.path.join(HERE, "..", "sqpsolver.jl_b200", "csrc", "symbolic.hpp"))):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    return C.CDLL(out)


def _solve(lib, J, P, d, w, rhs, tail_max=0, fuse=0):
    J = sp.csr_matrix(J); J.sort_indices()
    n, m = J.shape[1], J.shape[0]
    ip = C.POINTER(C.c_int32); dp = C.POINTER(C.c_double)
    a = lambda v, t: np.ascontiguousarray(v, dtype=t)
    jrp, jc, jv = a(J.indptr, np.int32), a(J.indices, np.int32), a(J.data, np.float64)
    if P is not None:
        P = sp.csr_matrix(P); P.sort_indices()
        prp, pc, pv = a(P.indptr, np.int32), a(P.indices, np.int32), a(P.data, np.float64)
    x = np.zeros(n); stats = np.zeros(6, dtype=np.int64)
    rc = lib.symcheck_solve3(n, m, jrp.ctypes.data_as(ip), jc.ctypes.data_as(ip), jv.ctypes.data_as(dp),
                            prp.ctypes.data_as(ip) if P is not None else None, pc.ctypes.data_as(ip) if P is not None else None,
                            pv.ctypes.data_as(dp) if P is not None else None, a(d, np.float64).ctypes.data_as(dp),
                            a(w, np.float64).ctypes.data_as(dp), a(rhs, np.float64).ctypes.data_as(dp), x.ctypes.data_as(dp),
                            stats.ctypes.data_as(C.POINTER(C.c_int64)), int(tail_max), int(fuse))
    return rc, x, stats


def test_random_patterns(symlib):
    rng = np.random.default_rng(0)
    for n, m, dens in ((8, 5, 0.4), (40, 60, 0.08), (120, 90, 0.03)):
        J = sp.random(m, n, dens, random_state=int(rng.integers(1 << 30)), data_rvs=rng.standard_normal).tocsr()
        M = sp.random(n, n, dens / 2, random_state=int(rng.integers(1 << 30)), data_rvs=rng.standard_normal)
        P = (M + M.T).tocsr()
        d = np.abs(P).sum(axis=1).A1 + rng.uniform(0.5, 2.0, n)  # diagonally dominant -> SPD
        w = rng.uniform(0.0, 3.0, m)
        w[::4] = 0.0  # inactive rows
        rhs = rng.standard_normal(n)
        K = P.toarray() + np.diag(d) + J.T.toarray() @ np.diag(w) @ J.toarray()
        for tail_max in (0, 24, 128):  # plain level-scheduled code / dense tail of the top levels
            for fuse in (0, 1):        # forward sweep on its own / riding in the factor phases
                rc, x, stats = _solve(symlib, J, P, d, w, rhs, tail_max, fuse)
                assert rc == 0, (n, m, tail_max, fuse, rc)
                assert np.abs(K @ x - rhs).max() <= 1e-10 * max(1.0, np.abs(rhs).max()), (n, m, tail_max, fuse)
                assert stats[4] <= tail_max


@pytest.mark.parametrize("fuse", [0, 1])
@pytest.mark.parametrize("tail_max", [0, 92])
@pytest.mark.parametrize("make", [lambda: AcopfPolar(case9()), lambda: AcopfPolar(synth_net(118, 186, 54, 118))])
def test_acopf_patterns_and_fill(symlib, make, tail_max, fuse):
    nlp = make()
    rng = np.random.default_rng(1)
    J = CooMatrix(nlp.j_row, nlp.j_col, nlp.m, nlp.n); J.fill(rng.standard_normal(nlp.nnz_jac_coo)); Js = J.to_scipy()
    H = SymCooMatrix(nlp.h_row, nlp.h_col, nlp.n); H.fill(0.01 * rng.standard_normal(nlp.nnz_hess_coo)); Hs = H.to_scipy()
    d = np.abs(Hs).sum(axis=1).A1 + 1e-3
    w = rng.uniform(0.0, 1e4, nlp.m)
    rhs = rng.standard_normal(nlp.n)
    rc, x, stats = _solve(symlib, Js, Hs, d, w, rhs, tail_max, fuse)
    assert rc == 0
    K = (Hs + sp.diags(d) + Js.T @ sp.diags(w) @ Js).toarray()
    ref = np.linalg.solve(K, rhs)
    assert np.abs(x - ref).max() <= 1e-8 * max(1.0, np.abs(ref).max())
    nnzL, nlev = int(stats[0]), int(stats[1])
    nnzK_lower = (np.count_nonzero(np.tril(K)))
    assert nnzL < 4 * nnzK_lower  # minimum degree keeps the fill small on network-structured problems
    assert nlev < nlp.n
    tail, tree_levels = int(stats[4]), int(stats[5])
    assert tail <= tail_max and nlev <= tree_levels
    if tail:  # the dense tail replaces the chain of narrow levels at the top of the elimination tree
        assert nlev < tree_levels


def test_negative_pivot_is_reported(symlib):
    J = sp.csr_matrix(np.array([[1.0, 1.0, 0.0]]))
    P = sp.csr_matrix(np.diag([1.0, -5.0, 1.0]))
    rc, x, _ = _solve(symlib, J, P, np.zeros(3), np.array([1.0]), np.ones(3))
    assert rc == -2


def _solve_ring(lib, J, P, d, w, rhs, tail_max, ns_max=512, kcap=4, stage_bytes=12320, stages=3):
    J = sp.csr_matrix(J); J.sort_indices()
    n, m = J.shape[1], J.shape[0]
    ip = C.POINTER(C.c_int32); dp = C.POINTER(C.c_double)
    a = lambda v, t: np.ascontiguousarray(v, dtype=t)
    jrp, jc, jv = a(J.indptr, np.int32), a(J.indices, np.int32), a(J.data, np.float64)
    if P is not None:
        P = sp.csr_matrix(P); P.sort_indices()
        prp, pc, pv = a(P.indptr, np.int32), a(P.indices, np.int32), a(P.data, np.float64)
    x = np.zeros(n); stats = np.zeros(4, dtype=np.int64)
    rc = lib.symcheck_ring(n, m, jrp.ctypes.data_as(ip), jc.ctypes.data_as(ip), jv.ctypes.data_as(dp),
                           prp.ctypes.data_as(ip) if P is not None else None, pc.ctypes.data_as(ip) if P is not None else None,
                           pv.ctypes.data_as(dp) if P is not None else None, a(d, np.float64).ctypes.data_as(dp),
                           a(w, np.float64).ctypes.data_as(dp), a(rhs, np.float64).ctypes.data_as(dp), x.ctypes.data_as(dp),
                           int(tail_max), int(ns_max), int(kcap), int(stage_bytes), int(stages), stats.ctypes.data_as(C.POINTER(C.c_int64)))
    return rc, x, stats


def test_ring_program_random_patterns(symlib):
    """The chunk images streamed into the shared-memory ring (build_ring_program) do the same factorisation and solves."""
    rng = np.random.default_rng(3)
    for n, m, dens in ((8, 5, 0.4), (40, 60, 0.08), (120, 90, 0.03), (300, 400, 0.02)):
        J = sp.random(m, n, dens, random_state=int(rng.integers(1 << 30)), data_rvs=rng.standard_normal).tocsr()
        M = sp.random(n, n, dens / 2, random_state=int(rng.integers(1 << 30)), data_rvs=rng.standard_normal)
        P = (M + M.T).tocsr()
        d = np.abs(P).sum(axis=1).A1 + rng.uniform(0.5, 2.0, n)
        w = rng.uniform(0.0, 3.0, m)
        w[::4] = 0.0
        rhs = rng.standard_normal(n)
        K = P.toarray() + np.diag(d) + J.T.toarray() @ np.diag(w) @ J.toarray()
        for tail_max in (0, 24, 96):
            for ns_max, kcap, stage in ((512, 4, 12320), (64, 2, 2048), (256, 8, 20000)):
                rc, x, stats = _solve_ring(symlib, J, P, d, w, rhs, tail_max, ns_max, kcap, stage)
                assert rc == 0, (n, m, tail_max, ns_max, rc)
                assert np.abs(K @ x - rhs).max() <= 1e-10 * max(1.0, np.abs(rhs).max()), (n, m, tail_max, ns_max)
                assert stats[2] * 4 <= stage
        rc, x, _ = _solve_ring(symlib, J, None, d + 1.0, w, rhs, 24)   # no P (the restoration LP)
        assert rc == 0
        K0 = np.diag(d + 1.0) + J.T.toarray() @ np.diag(w) @ J.toarray()
        assert np.abs(K0 @ x - rhs).max() <= 1e-10 * max(1.0, np.abs(rhs).max())


@pytest.mark.parametrize("make", [lambda: AcopfPolar(case9()), lambda: AcopfPolar(synth_net(118, 186, 54, 118))])
def test_ring_program_acopf(symlib, make):
    nlp = make()
    rng = np.random.default_rng(1)
    J = CooMatrix(nlp.j_row, nlp.j_col, nlp.m, nlp.n); J.fill(rng.standard_normal(nlp.nnz_jac_coo)); Js = J.to_scipy()
    H = SymCooMatrix(nlp.h_row, nlp.h_col, nlp.n); H.fill(0.01 * rng.standard_normal(nlp.nnz_hess_coo)); Hs = H.to_scipy()
    d = np.abs(Hs).sum(axis=1).A1 + 1e-3
    w = rng.uniform(0.0, 1e4, nlp.m)
    rhs = rng.standard_normal(nlp.n)
    rc, x, stats = _solve_ring(symlib, Js, Hs, d, w, rhs, 96)
    assert rc == 0
    K = (Hs + sp.diags(d) + Js.T @ sp.diags(w) @ Js).toarray()
    ref = np.linalg.solve(K, rhs)
    assert np.abs(x - ref).max() <= 1e-8 * max(1.0, np.abs(ref).max())
    print("ring program: chunks", stats[0], "words", stats[1], "largest image (bytes)", 4 * stats[2], "resident L entries", stats[3])
