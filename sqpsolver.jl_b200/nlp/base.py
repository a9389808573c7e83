"""NLP container handed to the SQP driver.

Mirrors the fields of the reference's ``Model{T,TD}`` (src/model.jl:3-67) that
the SQP trust-region driver reads: sizes, bounds, the 1-based COO structure of
the Jacobian (``j_str``) and of ONE triangle of the Lagrangian Hessian
(``h_str``), the five callbacks, ``num_linear_constraints`` and the start point.
The row ordering convention is the one ``MOI_wrapper.jl:759-766`` imposes:
linear <=, linear >=, linear ==, quadratic <=, >=, ==, then NLP rows.

Every callback is written so that ``x`` may carry leading batch dimensions
(``x.shape == (..., n)``): the batched workload (1024 perturbed-load ACOPF
instances that share one sparsity pattern) evaluates all instances in one
vectorised numpy call.
"""
from __future__ import annotations

import numpy as np


class NLP:
    """Base class; subclasses fill the attributes and override the callbacks."""

    name = "nlp"
    n: int
    m: int
    num_linear_constraints: int
    x_L: np.ndarray
    x_U: np.ndarray
    g_L: np.ndarray
    g_U: np.ndarray
    j_row: np.ndarray  # int64, 1-based, length nnzJ_coo (duplicates allowed)
    j_col: np.ndarray
    h_row: np.ndarray  # int64, 1-based, one triangle, duplicates allowed
    h_col: np.ndarray
    x0: np.ndarray
    sense = "Min"

    # -- callbacks (reference signatures: src/model.jl:20-25) ---------------
    def eval_f(self, x):
        raise NotImplementedError

    def eval_grad_f(self, x, grad):
        raise NotImplementedError

    def eval_g(self, x, g):
        raise NotImplementedError

    def eval_jac_g(self, x, values):
        raise NotImplementedError

    def eval_h(self, x, obj_factor, lam, values):
        raise NotImplementedError

    # -- helpers -------------------------------------------------------------
    @property
    def nnz_jac_coo(self):
        return int(self.j_row.shape[0])

    @property
    def nnz_hess_coo(self):
        return int(self.h_row.shape[0])

    def dense_jacobian(self, x):
        """Dense m x n Jacobian (tests only; duplicates summed in COO order)."""
        vals = np.empty(self.nnz_jac_coo)
        self.eval_jac_g(np.asarray(x, float), vals)
        J = np.zeros((self.m, self.n))
        for k in range(self.nnz_jac_coo):
            J[self.j_row[k] - 1, self.j_col[k] - 1] += vals[k]
        return J

    def dense_hessian(self, x, obj_factor, lam):
        """Dense symmetric Lagrangian Hessian, mirrored as sqp.jl:92-103 does."""
        vals = np.empty(self.nnz_hess_coo)
        self.eval_h(np.asarray(x, float), obj_factor, np.asarray(lam, float), vals)
        H = np.zeros((self.n, self.n))
        for k in range(self.nnz_hess_coo):
            r, c = self.h_row[k] - 1, self.h_col[k] - 1
            H[r, c] += vals[k]
            if r != c:
                H[c, r] += vals[k]
        return H
