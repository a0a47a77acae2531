"""DofHandler dof distribution (oracle; test infrastructure only).

Restates src/Dofs/DofHandler.jl of the reference for a single SubDofHandler that
covers the whole grid: `__close!` :493-569, `_close_subdofhandler!` :576-676,
`_distribute_dofs_for_cell!` :685-713, `add_vertex_dofs` :715-738,
`get_or_create_dofs!` :746-759, `add_face_dofs` :761-772, `add_edge_dofs` :774-785,
`add_volume_dofs` :787-794, `sortedge` :854-857, `sortface_fast` :1043-1057,
`dof_range` :1173-1187.  Orders 1 and 2 never permute entity dofs
(`adjust_dofs_during_distribution == false`, src/interpolations.jl:577-579).
"""
import numpy as np

__all__ = ["renumber_permutation", "renumber", "DofHandler"]


class DofHandler:
    def __init__(self, grid):
        self.grid = grid
        self.field_names = []
        self.field_ips = []
        self.closed = False
        self.cell_dofs = None      # (ncells, ndofs_per_cell) int64, 1-based
        self.ndofs = 0

    def add(self, name, ip):
        assert not self.closed
        assert ip.shape == self.grid.shape
        self.field_names.append(name)
        self.field_ips.append(ip)
        return self

    @property
    def ndofs_per_cell(self):
        return sum(ip.nbase for ip in self.field_ips)

    def field_offset(self, name):
        k = self.field_names.index(name)
        return sum(ip.nbase for ip in self.field_ips[:k])

    def dof_range(self, name):
        """1-based inclusive (first, last) like the reference's UnitRange."""
        off = self.field_offset(name)
        return (off + 1, off + self.field_ips[self.field_names.index(name)].nbase)

    def close(self):
        grid = self.grid
        rs = self.field_ips[0].refshape
        nf = len(self.field_ips)
        vertexdicts = [np.zeros(grid.nnodes + 1, dtype=np.int64) for _ in range(nf)]
        edgedicts = [dict() for _ in range(nf)]
        facedicts = [dict() for _ in range(nf)]
        infos = []
        for ip in self.field_ips:
            b = ip.base
            infos.append(dict(
                nv=[len(t) for t in b.vertexdof_indices],
                ne=[len(t) for t in b.edgedof_interior_indices],
                nfa=[len(t) for t in b.facedof_interior_indices],
                nvol=len(b.volumedof_interior_indices),
                ncopies=ip.vdim))
        nextdof = 1
        ndpc = self.ndofs_per_cell
        cells = grid.cells
        if nf == 1 and self.field_ips[0].order == 1 and grid.ncells > 20000:
            # Vertex-only single field: dofs are handed out in first-visit order of the (cell, local vertex)
            # sequence, vdim consecutive dofs per vertex -- the same result as the loop below, vectorised so that
            # the bench-sized baselines set up in seconds (cross-checked against the loop in the tests).
            nc = infos[0]["ncopies"]
            flat = cells.ravel()
            uniq, first = np.unique(flat, return_index=True)
            rank = np.empty(grid.nnodes + 1, dtype=np.int64)
            rank[uniq[np.argsort(first, kind="stable")]] = np.arange(len(uniq), dtype=np.int64)
            base = rank[cells] * nc + 1
            self.cell_dofs = (base[:, :, None] + np.arange(nc, dtype=np.int64)[None, None, :]).reshape(grid.ncells, ndpc)
            self.ndofs = len(uniq) * nc
            self.closed = True
            return self
        out = np.zeros((grid.ncells, ndpc), dtype=np.int64)
        for ci in range(grid.ncells):
            cell = cells[ci]
            col = 0
            row = out[ci]
            for f in range(nf):
                info = infos[f]
                nc = info["ncopies"]
                vd, ed, fd = vertexdicts[f], edgedicts[f], facedicts[f]
                # vertices
                for vi in range(rs.nvertices):
                    n = info["nv"][vi]
                    if n == 0:
                        continue
                    v = cell[vi]
                    first = vd[v]
                    if first == 0:
                        vd[v] = first = nextdof
                        nextdof += n * nc
                    for t in range(n * nc):
                        row[col] = first + t
                        col += 1
                # edges
                for ei, e in enumerate(rs.edges):
                    n = info["ne"][ei]
                    if n == 0:
                        continue
                    a, b = cell[e[0] - 1], cell[e[1] - 1]
                    key = (a, b) if a < b else (b, a)
                    first = ed.get(key)
                    if first is None:
                        ed[key] = first = nextdof
                        nextdof += n * nc
                    for t in range(n * nc):
                        row[col] = first + t
                        col += 1
                # faces
                for fi, fa in enumerate(rs.faces):
                    n = info["nfa"][fi]
                    if n == 0:
                        continue
                    key = tuple(sorted(cell[v - 1] for v in fa)[:3])
                    first = fd.get(key)
                    if first is None:
                        fd[key] = first = nextdof
                        nextdof += n * nc
                    for t in range(n * nc):
                        row[col] = first + t
                        col += 1
                # volume
                for t in range(info["nvol"] * nc):
                    row[col] = nextdof
                    nextdof += 1
                    col += 1
            assert col == ndpc
        self.cell_dofs = out
        self.ndofs = nextdof - 1
        self.closed = True
        return self

    def celldofs(self, ci):
        """1-based cell index -> dofs (1-based)."""
        return self.cell_dofs[ci - 1]


# ---- renumber! --------------------------------------------------------------------------------------------------
def renumber_permutation(dh, order, target_blocks=None):
    """compute_renumber_permutation (src/Dofs/DofRenumbering.jl:167-246): order 'fieldwise' | 'componentwise' with
    optional 1-based target blocks; returns perm (1-based: dof i becomes perm[i-1]).  Stable inside every block."""
    fdims = [ip.vdim for ip in dh.field_ips]
    ncomp = sum(fdims)
    if order == "fieldwise":
        tb = list(range(1, len(fdims) + 1)) if target_blocks is None else list(target_blocks)
        if len(tb) != len(fdims):
            raise ValueError("length of target block vector does not match number of fields and algebraic variables in DofHandler")
        comp_blocks = [tb[i] for i, d in enumerate(fdims) for _ in range(d)]
    elif order == "componentwise":
        comp_blocks = list(range(1, ncomp + 1)) if target_blocks is None else list(target_blocks)
        if len(comp_blocks) != ncomp:
            raise ValueError("length of target block vector does not match number of components in DofHandler")
    else:
        raise ValueError(order)
    if sorted(set(comp_blocks)) != list(range(1, max(comp_blocks) + 1)):
        raise ValueError("target blocks must be continuous and in the range 1:maxblock")
    nblocks = max(comp_blocks)
    blocks = [set() for _ in range(nblocks)]
    offs = np.concatenate([[0], np.cumsum(fdims)])
    col = 0
    for fi, ip in enumerate(dh.field_ips):
        nloc = ip.base.nbase * ip.vdim
        for j in range(nloc):
            comp = j % fdims[fi] + offs[fi]
            blocks[comp_blocks[comp] - 1].update(dh.cell_dofs[:, col + j].tolist())
        col += nloc
    iperm = []
    for b in blocks:
        iperm.extend(sorted(b))
    assert len(iperm) == dh.ndofs
    perm = np.empty(dh.ndofs, dtype=np.int64)
    perm[np.asarray(iperm) - 1] = np.arange(1, dh.ndofs + 1)
    return perm


def renumber(dh, perm, ch=None):
    """_renumber!(dh, perm) and _renumber!(ch, perm) (src/Dofs/DofRenumbering.jl:79-125)"""
    perm = np.asarray(perm, dtype=np.int64)
    assert sorted(perm.tolist()) == list(range(1, dh.ndofs + 1)), "input vector is not a permutation of length ndofs(dh)"
    dh.cell_dofs = perm[dh.cell_dofs - 1]
    if ch is not None:
        p = perm[np.asarray(ch.prescribed_dofs) - 1]
        o = np.argsort(p, kind="stable")
        ch.prescribed_dofs = p[o]
        ch.inhomogeneities = np.asarray(ch.inhomogeneities)[o]
    return dh
