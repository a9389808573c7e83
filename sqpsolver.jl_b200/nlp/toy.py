"""The reference's two toy NLPs, with analytic derivatives.

* :class:`ToyExample`  -- examples/toy_example.jl:17-23 (also test/ext_solver.jl:16-22):
  ``min X^2+X  s.t.  X^2-X == 2,  X*Y == 1,  X*Y >= 0,  X >= -2``.
  Row order follows MOI_wrapper.jl:759-766: the linear ``X >= -2`` row first,
  then the three NLP rows; ``num_linear_constraints == 1``.
  Pinned answer (test/runtests.jl:12-14): X ~ -1, Y ~ -1, LOCALLY_SOLVED.
* :class:`ReadmeToy` -- README.md:17-20: ``min x^2+x  s.t.  x^2-x == 2``
  (roots {-1, 2}; f(-1)=0 < f(2)=6).

The COO layouts are the ones MOI would hand to SqpSolver: affine rows first,
then the NLP block; the Hessian triangle lists the quadratic objective term
first (MOI_wrapper.jl:1010-1013) and then one entry per NLP-row term, so the
(1,1) slot has a duplicate on purpose.
"""
from __future__ import annotations

import numpy as np

from .base import NLP

INF = np.inf


class ToyExample(NLP):
    name = "toy_example"

    def __init__(self):
        self.n, self.m, self.num_linear_constraints = 2, 4, 1
        self.x_L = np.array([-INF, -INF])
        self.x_U = np.array([INF, INF])
        self.g_L = np.array([-2.0, 2.0, 1.0, 0.0])
        self.g_U = np.array([INF, 2.0, 1.0, INF])
        self.j_row = np.array([1, 2, 3, 3, 4, 4], dtype=np.int64)
        self.j_col = np.array([1, 1, 1, 2, 1, 2], dtype=np.int64)
        self.h_row = np.array([1, 1, 2, 2], dtype=np.int64)
        self.h_col = np.array([1, 1, 1, 1], dtype=np.int64)
        # default start: clamp(0, [l,u]) (MOI_wrapper.jl:1196-1197)
        self.x0 = np.zeros(2)

    def eval_f(self, x):
        X = x[..., 0]
        return X * X + X

    def eval_grad_f(self, x, grad):
        grad[..., 0] = 2.0 * x[..., 0] + 1.0
        grad[..., 1] = 0.0

    def eval_g(self, x, g):
        X, Y = x[..., 0], x[..., 1]
        g[..., 0] = X
        g[..., 1] = X * X - X
        g[..., 2] = X * Y
        g[..., 3] = X * Y

    def eval_jac_g(self, x, values):
        X, Y = x[..., 0], x[..., 1]
        values[..., 0] = 1.0
        values[..., 1] = 2.0 * X - 1.0
        values[..., 2] = Y
        values[..., 3] = X
        values[..., 4] = Y
        values[..., 5] = X

    def eval_h(self, x, obj_factor, lam, values):
        values[..., 0] = 2.0 * obj_factor
        values[..., 1] = 2.0 * lam[..., 1]
        values[..., 2] = lam[..., 2]
        values[..., 3] = lam[..., 3]


class ReadmeToy(NLP):
    name = "readme_toy"

    def __init__(self):
        self.n, self.m, self.num_linear_constraints = 1, 1, 0
        self.x_L = np.array([-INF])
        self.x_U = np.array([INF])
        self.g_L = np.array([2.0])
        self.g_U = np.array([2.0])
        self.j_row = np.array([1], dtype=np.int64)
        self.j_col = np.array([1], dtype=np.int64)
        self.h_row = np.array([1, 1], dtype=np.int64)
        self.h_col = np.array([1, 1], dtype=np.int64)
        self.x0 = np.zeros(1)

    def eval_f(self, x):
        X = x[..., 0]
        return X * X + X

    def eval_grad_f(self, x, grad):
        grad[..., 0] = 2.0 * x[..., 0] + 1.0

    def eval_g(self, x, g):
        X = x[..., 0]
        g[..., 0] = X * X - X

    def eval_jac_g(self, x, values):
        values[..., 0] = 2.0 * x[..., 0] - 1.0

    def eval_h(self, x, obj_factor, lam, values):
        values[..., 0] = 2.0 * obj_factor
        values[..., 1] = 2.0 * lam[..., 0]
