"""ConstraintHandler with Dirichlet conditions (oracle; test infrastructure only).

Restates src/Dofs/ConstraintHandler.jl of the reference:
  `Dirichlet` :41-51, `add!` :965-1008, `_add!` for facet/face/edge/vertex sets
  :404-428, `_local_facet_dofs_for_bc` :431-444, `_add!` for node sets :446-493,
  `add_prescribed_dof!` (later conditions override earlier ones) :383-401,
  `close!` (sort) :303-361, `update!`/`_update!` :504-580 with dof locations from
  `BCValues` (src/FEValues/FacetValues.jl:185-236), `apply!`/`apply_zero!` on (K, f)
  :710-740 (`meandiag` :952-958, `add_inhomogeneities_csc!` :755-768,
  `zero_out_columns!` :931-938, `zero_out_rows!` :940-950) and on vectors `_apply_v`
  :686-700.
"""
import numpy as np

from .interpolations import geometric_interpolation

__all__ = ["Dirichlet", "ConstraintHandler", "apply_local", "apply_assemble", "get_rhs_data", "apply_rhs"]


class Dirichlet:
    def __init__(self, field, entities, f, components=None, kind="facet"):
        """entities: (n,2) array of 1-based (cell, local entity) for kind in
        {'facet','face','edge','vertex'}, or a 1-D array of 1-based node ids for kind='node'.
        f(x, t) -> scalar or sequence with len(components) values."""
        self.field = field
        self.entities = np.asarray(entities, dtype=np.int64)
        self.f = f
        self.components = None if components is None else [int(c) for c in np.atleast_1d(components)]
        self.kind = kind


class ConstraintHandler:
    def __init__(self, dh):
        assert dh.closed
        self.dh = dh
        self.dbcs = []
        self._pre = []            # insertion-ordered prescribed dofs
        self._map = {}
        self.closed = False
        self.prescribed_dofs = None
        self.inhomogeneities = None

    def _add_dof(self, d):
        if d not in self._map:
            self._map[d] = len(self._pre)
            self._pre.append(d)

    def add(self, dbc):
        dh = self.dh
        ip = dh.field_ips[dh.field_names.index(dbc.field)]
        base, ncomp = ip.base, ip.vdim
        comps = list(range(1, ncomp + 1)) if not dbc.components else dbc.components
        assert comps == sorted(comps) and all(0 < c <= ncomp for c in comps)
        dbc._comps = comps
        offset = dh.field_offset(dbc.field)
        if dbc.kind == "node":
            nn = dh.grid.nnodes
            node_dofs = np.zeros((len(comps), nn + 1), dtype=np.int64)
            visited = np.zeros(nn + 1, dtype=bool)
            npts = min(base.nbase, dh.grid.cells.shape[1])
            for ci in range(dh.grid.ncells):
                for idx in range(npts):
                    node = dh.grid.cells[ci, idx]
                    if not visited[node]:
                        for i, c in enumerate(comps):
                            node_dofs[i, node] = dh.cell_dofs[ci, offset + idx * ncomp + c - 1]
                        visited[node] = True
            dbc._nodes, dbc._gdofs = [], []
            for node in dbc.entities:
                if not visited[node]:
                    continue
                dbc._nodes.append(int(node))
                for i in range(len(comps)):
                    dbc._gdofs.append(int(node_dofs[i, node]))
            for d in dbc._gdofs:
                self._add_dof(d)
        else:
            table = base.boundarydof_indices(dbc.kind)
            lfd = []      # per entity: list of local (0-based) cell-dof positions
            for ent in table:
                l = []
                for fdof in ent:
                    for d in range(1, ncomp + 1):
                        if d in comps:
                            l.append((fdof - 1) * ncomp + d + offset - 1)
                lfd.append(l)
            dbc._lfd = lfd
            # dof locations on each entity, in the geometric basis (BCValues)
            geo = geometric_interpolation(base.shape)
            dbc._M = [[geo.value_and_gradient(base.refcoords[fdof - 1])[0] for fdof in ent] for ent in table]
            for cell, ent in dbc.entities:
                for l in lfd[ent - 1]:
                    self._add_dof(int(dh.cell_dofs[cell - 1, l]))
        self.dbcs.append(dbc)
        return self

    def close(self, t=0.0):
        pre = np.array(self._pre, dtype=np.int64)
        order = np.argsort(pre, kind="stable")
        self.prescribed_dofs = pre[order]
        self._index = {int(d): i for i, d in enumerate(self.prescribed_dofs)}
        self.inhomogeneities = np.full(len(pre), np.nan)
        self.isconstrained = np.zeros(self.dh.ndofs + 1, dtype=bool)
        self.isconstrained[self.prescribed_dofs] = True
        self.closed = True
        self.update(t)
        return self

    def update(self, t=0.0):
        dh = self.dh
        for dbc in self.dbcs:
            nc = len(dbc._comps)
            if dbc.kind == "node":
                k = 0
                for node in dbc._nodes:
                    val = np.atleast_1d(dbc.f(dh.grid.nodes[node - 1], t))
                    assert len(val) == nc
                    for v in val:
                        self.inhomogeneities[self._index[dbc._gdofs[k]]] = v
                        k += 1
                continue
            for cell, ent in dbc.entities:
                coords = dh.grid.nodes[dh.grid.cells[cell - 1] - 1]
                l = dbc._lfd[ent - 1]
                counter = 0
                for M in dbc._M[ent - 1]:
                    x = np.zeros(coords.shape[1])
                    for i in range(len(M)):
                        x = x + M[i] * coords[i]
                    val = np.atleast_1d(dbc.f(x, t))
                    assert len(val) == nc
                    for i in range(nc):
                        g = int(dh.cell_dofs[cell - 1, l[counter]])
                        counter += 1
                        self.inhomogeneities[self._index[g]] = val[i]

    # ---- apply ---------------------------------------------------------------------
    def apply(self, K, f=None, applyzero=False):
        """apply!(K, f, ch) on an oracle CSC; returns the mean diagonal used."""
        n = K.n
        cp, rv, nz = K.colptr - 1, K.rowval - 1, K.nzval
        diag_pos = K.lookup(np.arange(1, n + 1), np.arange(1, n + 1))
        diag = np.where(diag_pos >= 0, nz[np.maximum(diag_pos, 0)], 0.0)
        m = 0.0
        for v in np.abs(diag):          # sequential sum like `meandiag`
            m += v
        m /= n
        pd = self.prescribed_dofs - 1
        if not applyzero and f is not None:
            for i, d in enumerate(pd):
                v = self.inhomogeneities[i]
                if v != 0:
                    r = slice(cp[d], cp[d + 1])
                    np.subtract.at(f, rv[r], v * nz[r])
        for d in pd:
            nz[cp[d]:cp[d + 1]] = 0.0
        nz[self.isconstrained[rv + 1]] = 0.0
        for i, d in enumerate(pd):
            nz[diag_pos[d]] = m
            if f is not None:
                f[d] = (0.0 if applyzero else self.inhomogeneities[i]) * m
        return m

    def apply_zero(self, K, f=None):
        return self.apply(K, f, applyzero=True)

    def apply_vec(self, u, applyzero=False):
        u[self.prescribed_dofs - 1] = 0.0 if applyzero else self.inhomogeneities
        return u


def apply_local(Ke, fe, dofs, ch, applyzero=False):
    """apply_local!(Ke, fe, global_dofs, ch; apply_zero) for Dirichlet constraints, in place
    (src/Dofs/ConstraintHandler.jl:1762-1822 `_apply_local!`, steps 1, 2 and 4; step 3 is the affine condensation).
    dofs 1-based."""
    n = len(dofs)
    index = {int(d): i for i, d in enumerate(ch.prescribed_dofs)}
    local = [(l, index[int(d)]) for l, d in enumerate(dofs) if int(d) in index]
    if not local:
        return
    for l, i in local:
        v = ch.inhomogeneities[i]
        if not applyzero and v != 0:
            for j in range(n):
                fe[j] -= v * Ke[j, l]
    m = 0.0
    for i in range(n):                  # meandiag, :952-958
        m += abs(Ke[i, i])
    m /= n
    for l, i in local:
        Ke[:, l] = 0.0
        Ke[l, :] = 0.0
        Ke[l, l] = m
        fe[l] = 0.0 if applyzero else ch.inhomogeneities[i] * m


def apply_assemble(K, f, ch, dofs, Ke, fe, applyzero=False):
    """apply_assemble!(assembler, ch, global_dofs, Ke, fe; apply_zero), src/assembler.jl:491-503 (destructive on Ke, fe)"""
    from .assemble import assemble_cell
    apply_local(Ke, fe, dofs, ch, applyzero)
    assemble_cell(K, f, dofs, Ke, fe)


def get_rhs_data(ch, K):
    """get_rhs_data(ch, A) (src/Dofs/ConstraintHandler.jl:203-208): (meandiag(A), A[:, prescribed_dofs]) of the matrix as it is"""
    n = K.n
    cp = K.colptr - 1
    diag_pos = K.lookup(np.arange(1, n + 1), np.arange(1, n + 1))
    m = 0.0
    for v in np.abs(np.where(diag_pos >= 0, K.nzval[np.maximum(diag_pos, 0)], 0.0)):
        m += v
    m /= n
    cols = [(K.rowval[cp[d]:cp[d + 1]] - 1, K.nzval[cp[d]:cp[d + 1]].copy()) for d in ch.prescribed_dofs - 1]
    return m, cols


def apply_rhs(data, f, ch, applyzero=False):
    """apply_rhs!(data, f, ch, applyzero) (:217-240) without the affine branch"""
    m, cols = data
    for i, (rows, vals) in enumerate(cols):
        v = ch.inhomogeneities[i]
        if not applyzero and v != 0:
            np.subtract.at(f, rows, v * vals)
    for i, d in enumerate(ch.prescribed_dofs - 1):
        f[d] = (0.0 if applyzero else ch.inhomogeneities[i]) * m
    return f
