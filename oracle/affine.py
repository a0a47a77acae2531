"""Affine constraints (oracle; test infrastructure only) -- groundwork for SURVEY.md 8f-3, NOT yet built in the CUDA library.

Restates, for the next round's device implementation:
  * `AffineConstraint(dof, [master => coeff, ...], b)` and `add!` / `close!` (src/Dofs/ConstraintHandler.jl:114-131, 303-361:
    sorted prescribed dofs, `dofcoefficients` aligned with them, nested constraints rejected),
  * `create_constraint_matrix` (:889-916): a = C a_f + g,
  * the condensed sparsity pattern of `allocate_matrix(dh, ch)` (src/Dofs/sparsity_pattern.jl:782-840 `_add_constraint_entries!`:
    entry (r, c) with r constrained adds (m, c) for the masters m of r, c constrained adds (r, m), both add (m1, m2)),
  * `apply!(K, f, ch)` with `_condense!` (:710-740, 809-868) and `apply!(u, ch)` (:686-700).
Restriction of this restatement: master dofs must be unconstrained (the reference additionally allows Dirichlet-prescribed
masters through `affine_inhomogeneities`, :338-345).
Pinned by tests/test_oracle_goldens.py::test_affine_constraints_golden on the identities of test/test_constraints.jl:323-412.
"""
import numpy as np

from .constraints import ConstraintHandler
from .pattern import CSC, allocate_matrix

__all__ = ["AffineConstraint", "AffineConstraintHandler", "allocate_matrix_condensed", "dense_pattern"]


class AffineConstraint:
    def __init__(self, constrained_dof, entries, b):
        self.constrained_dof = int(constrained_dof)
        self.entries = [(int(d), float(v)) for d, v in entries]
        self.b = float(b)


class AffineConstraintHandler(ConstraintHandler):
    def __init__(self, dh):
        super().__init__(dh)
        self._affine = {}

    def add(self, c):
        if isinstance(c, AffineConstraint):
            assert c.constrained_dof not in self._map, "oracle: overwriting an existing constraint is out of scope"
            self._add_dof(c.constrained_dof)
            self._affine[c.constrained_dof] = c
            return self
        return super().add(c)

    def close(self, t=0.0):
        super().close(t)
        self.dofcoefficients = [self._affine[int(d)].entries if int(d) in self._affine else None for d in self.prescribed_dofs]
        for coeffs in self.dofcoefficients:
            for d, _ in coeffs or []:
                if d in self._index:
                    if self.dofcoefficients[self._index[d]]:
                        raise ValueError("nested affine constraints currently not supported")
                    raise NotImplementedError("oracle: master dofs prescribed by Dirichlet conditions are out of scope")
        self.free_dofs = np.setdiff1d(np.arange(1, self.dh.ndofs + 1), self.prescribed_dofs)
        return self

    def update(self, t=0.0):
        super().update(t)
        for d, c in self._affine.items():
            self.inhomogeneities[self._index[d]] = c.b

    def create_constraint_matrix(self):
        import scipy.sparse as sp
        n, nf = self.dh.ndofs, len(self.free_dofs)
        I, J, V = list(self.free_dofs - 1), list(range(nf)), [1.0] * nf
        for i, pdof in enumerate(self.prescribed_dofs):
            for d, v in self.dofcoefficients[i] or []:
                I.append(int(pdof) - 1)
                J.append(int(np.searchsorted(self.free_dofs, d)))
                V.append(v)
        g = np.zeros(n)
        g[self.prescribed_dofs - 1] = self.inhomogeneities
        return sp.csc_matrix((V, (I, J)), shape=(n, nf)), g

    def _coeffs(self, dof):
        i = self._index.get(int(dof))
        return None if i is None else self.dofcoefficients[i]

    def apply(self, K, f=None, applyzero=False):
        """apply!(K, f, ch, applyzero) with affine constraints; returns the mean diagonal."""
        n = K.n
        cp, rv, nz = K.colptr - 1, K.rowval - 1, K.nzval
        diag_pos = K.lookup(np.arange(1, n + 1), np.arange(1, n + 1))
        m = 0.0
        for v in np.abs(np.where(diag_pos >= 0, nz[np.maximum(diag_pos, 0)], 0.0)):
            m += v
        m /= n
        if not applyzero and f is not None:            # add_inhomogeneities!: f -= K * g
            for i, d in enumerate(self.prescribed_dofs - 1):
                v = self.inhomogeneities[i]
                if v != 0:
                    r = slice(cp[d], cp[d + 1])
                    np.subtract.at(f, rv[r], v * nz[r])
        if any(self.dofcoefficients[i] for i in range(len(self.prescribed_dofs))):      # _condense!, :809-868
            def addindex(val, row, col):
                p = K.lookup(np.array([row]), np.array([col]))[0]
                if p < 0:
                    raise KeyError(f"condensation needs entry ({row}, {col}): use allocate_matrix_condensed")
                nz[p] += val
            for col in range(1, n + 1):
                ccol = self._coeffs(col)
                for a in range(cp[col - 1], cp[col]):
                    kv = nz[a]
                    if kv == 0:
                        continue
                    row = rv[a] + 1
                    crow = self._coeffs(row)
                    if ccol is None:
                        for d, v in crow or []:
                            addindex(v * kv, d, col)
                    elif crow is None:
                        for d, v in ccol:
                            addindex(v * kv, row, d)
                    else:
                        for d1, v1 in crow:
                            for d2, v2 in ccol:
                                addindex(v1 * v2 * kv, d1, d2)
                if ccol is not None and f is not None:
                    for d, v in ccol:
                        f[d - 1] += f[col - 1] * v
                    f[col - 1] = 0.0
        for d in self.prescribed_dofs - 1:
            nz[cp[d]:cp[d + 1]] = 0.0
        nz[self.isconstrained[rv + 1]] = 0.0
        for i, d in enumerate(self.prescribed_dofs - 1):
            nz[diag_pos[d]] = m
            if f is not None:
                f[d] = (0.0 if applyzero else self.inhomogeneities[i]) * m
        return m

    def apply_vec(self, u, applyzero=False):
        u[self.prescribed_dofs - 1] = 0.0 if applyzero else self.inhomogeneities
        for i, dof in enumerate(self.prescribed_dofs):
            coeffs = self.dofcoefficients[i]
            if coeffs is None:
                continue
            u[dof - 1] = 0.0 if applyzero else self.inhomogeneities[i]
            for d, s in coeffs:
                u[dof - 1] += s * u[d - 1]
        return u


def _csc_from_pairs(n, pairs):
    keys = np.unique(np.array([(c - 1) * n + (r - 1) for r, c in pairs], dtype=np.int64))
    cols, rows = keys // n, keys % n
    colptr = np.concatenate([[0], np.cumsum(np.bincount(cols, minlength=n))]) + 1
    return CSC(n, colptr, rows + 1)


def allocate_matrix_condensed(dh, ch):
    """allocate_matrix(dh, ch): the pattern of allocate_matrix(dh) plus the entries `_condense!` writes"""
    K = allocate_matrix(dh)
    n = K.n
    cols = np.repeat(np.arange(1, n + 1), np.diff(K.colptr))
    pairs = set(zip(K.rowval.tolist(), cols.tolist()))
    extra = set()
    for r, c in pairs:
        cr, cc = ch._coeffs(r), ch._coeffs(c)      # None: not affinely constrained (free or Dirichlet), sparsity_pattern.jl:794-832
        if cr is None and cc is None:
            continue
        if cr is None:
            extra.update((r, d) for d, _ in cc)
        elif cc is None:
            extra.update((d, c) for d, _ in cr)
        else:
            extra.update((d1, d2) for d1, _ in cr for d2, _ in cc)
    return _csc_from_pairs(n, pairs | extra)


def dense_pattern(n):
    return CSC(n, np.arange(0, n * n + 1, n) + 1, np.tile(np.arange(1, n + 1), n))
