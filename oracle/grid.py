"""Grid + structured generators (oracle; test infrastructure only).

Restates src/Grid/grid_generators.jl of the reference: Line :8-37, Quadrilateral
:78-112, Triangle :383-417, Hexahedron :159-203, Tetrahedron :474-537, node
generator `_generate_nodes` :550-578.  Cells hold 1-based node ids; facetsets are
arrays of (cell, local facet), 1-based, sorted by (cell, facet) like the
reference's `sort!(s, by = x -> x.idx)`.
"""
import numpy as np

from .interpolations import Lagrange

__all__ = ["Grid", "generate_grid", "perturb_grid"]

_CELLSHAPE_NNODES = {"line": 2, "triangle": 3, "quadrilateral": 4, "tetrahedron": 4, "hexahedron": 8}


class Grid:
    def __init__(self, shape, cells, nodes, facetsets=None):
        self.shape = shape
        self.cells = np.ascontiguousarray(cells, dtype=np.int64)      # (ncells, nnpc), 1-based
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)    # (nnodes, sdim)
        self.facetsets = facetsets or {}
        assert self.cells.shape[1] == _CELLSHAPE_NNODES[shape]

    @property
    def ncells(self):
        return self.cells.shape[0]

    @property
    def nnodes(self):
        return self.nodes.shape[0]

    @property
    def sdim(self):
        return self.nodes.shape[1]


def _generate_nodes(shape, nn, corners):
    """src/Grid/grid_generators.jl:550-559: xi = 2(idx-1)/(nn-1) - 1, x = sum_i M_i(xi) corner_i,
    first index fastest."""
    ip = Lagrange(shape, 1)
    rdim = len(nn)
    corners = np.asarray(corners, dtype=np.float64)
    axes = [2.0 * np.arange(n, dtype=np.float64) / (n - 1) - 1.0 for n in nn]
    mesh = np.meshgrid(*axes, indexing="ij")
    xi = np.stack([m.ravel(order="F") for m in mesh], axis=1)      # (nnodes, rdim), first index fastest
    # M_i(xi) for the linear hypercube: prod_d (1 +- xi_d)/2^rdim, evaluated like the reference formula
    x = np.zeros((xi.shape[0], corners.shape[1]))
    for i in range(ip.nbase):
        Mi = np.ones(xi.shape[0])
        for d in range(rdim):
            s = ip.refcoords[i, d]
            Mi = Mi * ((1 + xi[:, d]) if s > 0 else (1 - xi[:, d]))
        Mi = Mi / (2 ** rdim)
        x += Mi[:, None] * corners[i][None, :]
    return x


def _extrema_to_corners(shape, left, right):
    """src/Grid/grid_generators.jl:565-578"""
    ip = Lagrange(shape, 1)
    left = np.asarray(left, dtype=np.float64)
    right = np.asarray(right, dtype=np.float64)
    dx = right - left
    out = []
    for xi in ip.refcoords:
        dxi = xi - (-1.0)
        out.append(left + (dx / 2.0) * dxi)
    return np.array(out)


def _sorted_set(pairs):
    a = np.array(sorted(set(map(tuple, pairs))), dtype=np.int64).reshape(-1, 2)
    return a


def generate_grid(shape, nel, left=None, right=None):
    nel = tuple(int(n) for n in nel)
    dim = len(nel)
    if left is None:
        left = (-1.0,) * dim
    if right is None:
        right = (1.0,) * dim
    if shape == "line":
        (nx,) = nel
        nodes = _generate_nodes("line", (nx + 1,), _extrema_to_corners("line", left, right))
        cells = np.array([(i, i + 1) for i in range(1, nx + 1)], dtype=np.int64)
        fs = {"left": _sorted_set([(1, 1)]), "right": _sorted_set([(nx, 2)])}
        return Grid("line", cells, nodes, fs)

    if shape in ("quadrilateral", "triangle"):
        nx, ny = nel
        corners = _extrema_to_corners("quadrilateral", left, right)
        nodes = _generate_nodes("quadrilateral", (nx + 1, ny + 1), corners)
        na = np.arange(1, (nx + 1) * (ny + 1) + 1, dtype=np.int64).reshape((nx + 1, ny + 1), order="F")
        cells = []
        if shape == "quadrilateral":
            for j in range(ny):
                for i in range(nx):
                    cells.append((na[i, j], na[i + 1, j], na[i + 1, j + 1], na[i, j + 1]))
            ca = np.arange(1, nx * ny + 1, dtype=np.int64).reshape((nx, ny), order="F")
            fs = {
                "bottom": _sorted_set([(c, 1) for c in ca[:, 0]]),
                "right": _sorted_set([(c, 2) for c in ca[-1, :]]),
                "top": _sorted_set([(c, 3) for c in ca[:, -1]]),
                "left": _sorted_set([(c, 4) for c in ca[0, :]]),
            }
        else:
            for j in range(ny):
                for i in range(nx):
                    cells.append((na[i, j], na[i + 1, j], na[i, j + 1]))
                    cells.append((na[i + 1, j], na[i + 1, j + 1], na[i, j + 1]))
            ca = np.arange(1, 2 * nx * ny + 1, dtype=np.int64).reshape((2, nx, ny), order="F")
            fs = {
                "bottom": _sorted_set([(c, 1) for c in ca[0, :, 0]]),
                "right": _sorted_set([(c, 1) for c in ca[1, -1, :]]),
                "top": _sorted_set([(c, 2) for c in ca[1, :, -1]]),
                "left": _sorted_set([(c, 3) for c in ca[0, 0, :]]),
            }
        return Grid(shape, np.array(cells, dtype=np.int64), nodes, fs)

    if shape in ("hexahedron", "tetrahedron"):
        nx, ny, nz = nel
        corners = _extrema_to_corners("hexahedron", left, right)
        nodes = _generate_nodes("hexahedron", (nx + 1, ny + 1, nz + 1), corners)
        na = np.arange(1, (nx + 1) * (ny + 1) * (nz + 1) + 1, dtype=np.int64).reshape(
            (nx + 1, ny + 1, nz + 1), order="F")
        i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
        i, j, k = (a.ravel(order="F") for a in (i, j, k))         # i fastest
        c = np.stack([na[i, j, k], na[i + 1, j, k], na[i + 1, j + 1, k], na[i, j + 1, k],
                      na[i, j, k + 1], na[i + 1, j, k + 1], na[i + 1, j + 1, k + 1], na[i, j + 1, k + 1]], axis=1)
        if shape == "hexahedron":
            ca = np.arange(1, nx * ny * nz + 1, dtype=np.int64).reshape((nx, ny, nz), order="F")
            fs = {
                "bottom": _sorted_set([(x, 1) for x in ca[:, :, 0].ravel()]),
                "front": _sorted_set([(x, 2) for x in ca[:, 0, :].ravel()]),
                "right": _sorted_set([(x, 3) for x in ca[-1, :, :].ravel()]),
                "back": _sorted_set([(x, 4) for x in ca[:, -1, :].ravel()]),
                "left": _sorted_set([(x, 5) for x in ca[0, :, :].ravel()]),
                "top": _sorted_set([(x, 6) for x in ca[:, :, -1].ravel()]),
            }
            return Grid("hexahedron", c, nodes, fs)
        # 6 tets per cube: src/Grid/grid_generators.jl:506-511 (1-based corner ids of the cube)
        split = [(1, 2, 4, 8), (1, 5, 2, 8), (2, 3, 4, 8), (2, 7, 3, 8), (2, 5, 6, 8), (2, 6, 7, 8)]
        cells = np.empty((6 * c.shape[0], 4), dtype=np.int64)
        for t, s in enumerate(split):
            cells[t::6, :] = c[:, [v - 1 for v in s]]
        cn = np.arange(1, 6 * nx * ny * nz + 1, dtype=np.int64).reshape((6, nx, ny, nz), order="F")

        def fl(a):
            return a.ravel()
        fs = {
            "left": _sorted_set([(x, 4) for x in fl(cn[0, 0, :, :])] + [(x, 2) for x in fl(cn[1, 0, :, :])]),
            "right": _sorted_set([(x, 1) for x in fl(cn[3, -1, :, :])] + [(x, 1) for x in fl(cn[5, -1, :, :])]),
            "front": _sorted_set([(x, 1) for x in fl(cn[1, :, 0, :])] + [(x, 1) for x in fl(cn[4, :, 0, :])]),
            "back": _sorted_set([(x, 3) for x in fl(cn[2, :, -1, :])] + [(x, 3) for x in fl(cn[3, :, -1, :])]),
            "bottom": _sorted_set([(x, 1) for x in fl(cn[0, :, :, 0])] + [(x, 1) for x in fl(cn[2, :, :, 0])]),
            "top": _sorted_set([(x, 3) for x in fl(cn[4, :, :, -1])] + [(x, 3) for x in fl(cn[5, :, :, -1])]),
        }
        return Grid("tetrahedron", cells, nodes, fs)
    raise ValueError(shape)


def _hash01(ids, salt):
    """Deterministic integer hash -> [0,1).  Same formula as fb2_grid_perturb in the CUDA
    library (ferrite.jl_b200/csrc/host_grid.cpp): splitmix64 of (id*4 + salt)."""
    x = (ids.astype(np.uint64) * np.uint64(4) + np.uint64(salt)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def perturb_grid(grid, nel, left, right, amplitude=0.2):
    """Deterministic interior-node perturbation x += amplitude*h*(hash-1/2) so that cells are
    not all congruent (cf. the reference's `perturb_standard_grid!`, test/test_utils.jl:282).
    Boundary nodes of the generated box stay fixed.  Node ids are 1-based in the hash."""
    nel = np.asarray(nel)
    left = np.asarray(left, dtype=np.float64)
    right = np.asarray(right, dtype=np.float64)
    dim = len(nel)
    nn = nel + 1
    ids = np.arange(grid.nnodes, dtype=np.int64)
    idx = []
    rem = ids.copy()
    for d in range(dim):
        idx.append(rem % nn[d])
        rem //= nn[d]
    interior = np.ones(grid.nnodes, dtype=bool)
    for d in range(dim):
        interior &= (idx[d] > 0) & (idx[d] < nn[d] - 1)
    h = (right - left) / nel
    for d in range(dim):
        r = _hash01(ids + 1, d)
        grid.nodes[:, d] += np.where(interior, amplitude * h[d] * (r - 0.5), 0.0)
    return grid
