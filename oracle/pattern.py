"""Sparsity pattern -> CSC (oracle; test infrastructure only).

Restates the *generic* builder of the reference, the simplest specification of
the result: every cell couples all its dofs with each other, the diagonal is
always present (src/Dofs/sparsity_pattern.jl:370-398, 742-773, 1136-1204), and
the CSC has ascending unique rows within each column (:916-945, 951-991).
Arrays are 1-based Int64 like `SparseMatrixCSC{Float64,Int}`.
"""
import numpy as np

__all__ = ["allocate_matrix", "CSC"]


class CSC:
    def __init__(self, n, colptr, rowval, nzval=None):
        self.n = int(n)
        self.colptr = np.asarray(colptr, dtype=np.int64)     # 1-based, length n+1
        self.rowval = np.asarray(rowval, dtype=np.int64)     # 1-based
        self.nzval = np.zeros(len(self.rowval)) if nzval is None else np.asarray(nzval, dtype=np.float64)

    @property
    def nnz(self):
        return len(self.rowval)

    def toscipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.nzval, self.rowval - 1, self.colptr - 1), shape=(self.n, self.n))

    def lookup(self, rows, cols):
        """Position (0-based, into nzval) of entries (rows, cols), both 1-based arrays; -1 if absent."""
        keys = (self.colkeys())
        q = (np.asarray(cols, dtype=np.int64) - 1) * self.n + (np.asarray(rows, dtype=np.int64) - 1)
        pos = np.searchsorted(keys, q)
        pos_c = np.minimum(pos, len(keys) - 1)
        ok = keys[pos_c] == q
        return np.where(ok, pos_c, -1)

    def colkeys(self):
        if not hasattr(self, "_keys"):
            cols = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(self.colptr))
            self._keys = cols * self.n + (self.rowval - 1)
        return self._keys


def allocate_matrix(dh, chunk=1 << 22):
    """`allocate_matrix(dh)`: union over cells of dofs x dofs, plus the diagonal."""
    n = dh.ndofs
    cd = dh.cell_dofs - 1
    ncells, ndpc = cd.shape
    parts = [np.arange(n, dtype=np.int64) * n + np.arange(n, dtype=np.int64)]   # diagonal
    per = max(1, chunk // (ndpc * ndpc))
    for s in range(0, ncells, per):
        c = cd[s:s + per]
        keys = (c[:, None, :] * n + c[:, :, None]).ravel()   # col*n + row
        parts.append(np.unique(keys))
    uniq = np.unique(np.concatenate(parts))
    cols = uniq // n
    rows = uniq % n
    colptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(colptr, cols + 1, 1)
    colptr = np.cumsum(colptr) + 1
    K = CSC(n, colptr, rows + 1)
    K._keys = uniq
    return K
