"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's COO -> sparse assembly.

Follows
  * pattern:  ``sparse(j_row, j_col, ones, m, n)`` sqp_trust_region.jl:47-48 and
              ``sparse(h_row, h_col, ones, n, n)`` :56-57 (Julia CSC: column-major,
              row indices ascending inside a column, duplicates merged);
  * values:   ``fill!(nzval, 0); Jacobian[r,c] += v`` for k ascending, sqp.jl:111-117;
              Hessian: diagonal added once, off-diagonal added to (r,c) AND (c,r),
              sqp.jl:92-103.
Each slot therefore holds ``((0.0 + v_k1) + v_k2) + ...`` with k1 < k2 < ... the
COO positions that map to it -- a strictly sequential fp64 sum, reproduced here
without any pairwise/blocked summation so the device scatter can be compared
bit-for-bit.

The device stores J as CSR (row-major) plus the CSR of J^T; the CSR of J^T *is*
Julia's CSC of J (same colptr / rowval / nzval order), which is what
:func:`csc_pattern` returns (0-based).
"""
from __future__ import annotations

import numpy as np


def _pattern(major, minor, n_major):
    """Sorted-unique pattern of (major, minor) pairs; returns ptr, idx, slot_of_entry."""
    nnz = major.shape[0]
    order = np.lexsort((np.arange(nnz), minor, major))  # by major, then minor, then k (stable)
    smaj, smin = major[order], minor[order]
    head = np.ones(nnz, dtype=bool)
    head[1:] = (smaj[1:] != smaj[:-1]) | (smin[1:] != smin[:-1])
    slot_sorted = np.cumsum(head) - 1
    slot = np.empty(nnz, dtype=np.int64)
    slot[order] = slot_sorted
    idx = smin[head].astype(np.int32)
    counts = np.bincount(smaj[head], minlength=n_major)
    ptr = np.zeros(n_major + 1, dtype=np.int32)
    np.cumsum(counts, out=ptr[1:])
    return ptr, idx, slot, order


def csr_pattern(row1, col1, nrows):
    """CSR pattern from 1-based COO.  Returns (row_ptr, col_idx, slot[k])."""
    ptr, idx, slot, _ = _pattern(np.asarray(row1) - 1, np.asarray(col1) - 1, nrows)
    return ptr, idx, slot


def csc_pattern(row1, col1, ncols):
    """Julia's ``sparse(I,J,V,m,n)`` layout, 0-based: (colptr, rowval, slot[k])."""
    ptr, idx, slot, _ = _pattern(np.asarray(col1) - 1, np.asarray(row1) - 1, ncols)
    return ptr, idx, slot


def sym_expand(h_row1, h_col1):
    """Entry list of the symmetric fill of sqp.jl:92-103.

    Returns (rows0, cols0, src) where entry e adds ``h_val[src[e]]`` to slot
    (rows0[e], cols0[e]); for an off-diagonal COO entry k the (r,c) contribution
    is listed before the (c,r) one, both tagged with source k, so a stable sort
    by (slot, src) reproduces Julia's accumulation order.
    """
    r = np.asarray(h_row1) - 1
    c = np.asarray(h_col1) - 1
    k = np.arange(r.shape[0])
    off = r != c
    rows = np.concatenate([r, c[off]])
    cols = np.concatenate([c, r[off]])
    src = np.concatenate([k, k[off]])
    return rows, cols, src


def ordered_scatter(slot, src, vals, nslots):
    """out[s] = sequential sum (ascending src) of vals[src[e]] over entries with slot[e]==s."""
    order = np.lexsort((src, slot))
    s_sorted = slot[order]
    v_sorted = np.asarray(vals, float)[src[order]]
    counts = np.bincount(s_sorted, minlength=nslots)
    start = np.zeros(nslots + 1, dtype=np.int64)
    np.cumsum(counts, out=start[1:])
    out = np.zeros(nslots)
    for j in range(int(counts.max()) if nslots else 0):
        sel = counts > j
        out[sel] = out[sel] + v_sorted[start[:-1][sel] + j]
    return out


class CooMatrix:
    """Jacobian-like matrix: fixed 1-based COO pattern, values re-scattered per iterate."""

    def __init__(self, row1, col1, nrows, ncols):
        self.shape = (nrows, ncols)
        self.row_ptr, self.col_idx, self.slot = csr_pattern(row1, col1, nrows)
        self.nnz = int(self.col_idx.shape[0])
        self.src = np.arange(len(row1))
        self.data = np.zeros(self.nnz)

    def fill(self, vals):
        self.data = ordered_scatter(self.slot, self.src, vals, self.nnz)
        return self.data

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csr_matrix((self.data, self.col_idx, self.row_ptr), shape=self.shape)


class SymCooMatrix(CooMatrix):
    """Hessian: one-triangle COO in, symmetric-full CSR out (sqp.jl:92-103)."""

    def __init__(self, h_row1, h_col1, n):
        rows, cols, src = sym_expand(h_row1, h_col1)
        self.shape = (n, n)
        self.row_ptr, self.col_idx, self.slot = csr_pattern(rows + 1, cols + 1, n)
        self.nnz = int(self.col_idx.shape[0])
        self.src = src
        self.data = np.zeros(self.nnz)
