"""Lagrange interpolations, orders 1 and 2 (oracle; test infrastructure only).

Restates src/interpolations.jl of the reference:
  * reference coordinates + shape functions: Line :600-635, Quadrilateral :640-696,
    Triangle :738-790, Tetrahedron :891-952, Hexahedron :1038-1158
  * entity dof index tables (vertex/edge/face/volume interior): :584-590, :667-668,
    :766, :918, :1075-1083; `edgedof_indices`/`facedof_indices` :349-365, :402-422
  * VectorizedInterpolation dof interleave `(i-1)*vdim + c` :1816-1825
Shape-function gradients: the reference differentiates the formulas with
ForwardDiff (src/interpolations.jl:280-292); here the analytic derivatives of the
same formulas are used (equal up to rounding).
"""
import numpy as np

from .refshapes import REFSHAPES

__all__ = ["Lagrange", "VectorLagrange", "geometric_interpolation"]

_REFCOORDS = {
    ("line", 1): [(-1.0,), (1.0,)],
    ("line", 2): [(-1.0,), (1.0,), (0.0,)],
    ("quadrilateral", 1): [(-1, -1), (1, -1), (1, 1), (-1, 1)],
    ("quadrilateral", 2): [(-1, -1), (1, -1), (1, 1), (-1, 1), (0, -1), (1, 0), (0, 1), (-1, 0), (0, 0)],
    ("hexahedron", 1): [(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1),
                        (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)],
    ("hexahedron", 2): [
        (-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1),
        (0, -1, -1), (1, 0, -1), (0, 1, -1), (-1, 0, -1), (0, -1, 1), (1, 0, 1), (0, 1, 1), (-1, 0, 1),
        (-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0),
        (0, 0, -1), (0, -1, 0), (1, 0, 0), (0, 1, 0), (-1, 0, 0), (0, 0, 1),
        (0, 0, 0)],
    ("triangle", 1): [(1, 0), (0, 1), (0, 0)],
    ("triangle", 2): [(1, 0), (0, 1), (0, 0), (0.5, 0.5), (0, 0.5), (0.5, 0)],
    ("tetrahedron", 1): [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)],
    ("tetrahedron", 2): [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0.5, 0, 0), (0.5, 0.5, 0),
                         (0, 0.5, 0), (0, 0, 0.5), (0.5, 0, 0.5), (0, 0.5, 0.5)],
}


def _rng(a, b):
    return tuple((i,) for i in range(a, b + 1))


# (edge interior, face interior, volume interior), 1-based
_INTERIOR = {
    ("line", 1): (((),), (), ()),
    ("line", 2): (((3,),), (), ()),
    ("quadrilateral", 1): (((),) * 4, ((),), ()),
    ("quadrilateral", 2): (_rng(5, 8), ((9,),), ()),
    ("triangle", 1): (((),) * 3, ((),), ()),
    ("triangle", 2): (_rng(4, 6), ((),), ()),
    ("tetrahedron", 1): (((),) * 6, ((),) * 4, ()),
    ("tetrahedron", 2): (_rng(5, 10), ((),) * 4, ()),
    ("hexahedron", 1): (((),) * 12, ((),) * 6, ()),
    ("hexahedron", 2): (_rng(9, 20), _rng(21, 26), (27,)),
}


def _l1d(order, node, x):
    """1-D Lagrange factor at reference coordinate `node` in {-1,0,1}; returns (value, derivative)."""
    if order == 1:
        if node < 0:
            return (1 - x) / 2, -0.5 + 0 * x
        return (1 + x) / 2, 0.5 + 0 * x
    if node < 0:   # phi1 = -x(1-x)/2
        return -x * (1 - x) / 2, x - 0.5
    if node > 0:   # phi3 = x(1+x)/2
        return x * (1 + x) / 2, x + 0.5
    return (1 + x) * (1 - x), -2 * x   # phi2


class Lagrange:
    """Scalar Lagrange{shape, order}, order in (1, 2)."""

    def __init__(self, shape, order):
        assert (shape, order) in _REFCOORDS, f"Lagrange{{{shape},{order}}} not in oracle scope"
        self.shape = shape
        self.order = order
        self.refshape = REFSHAPES[shape]
        self.rdim = self.refshape.rdim
        self.refcoords = np.array(_REFCOORDS[(shape, order)], dtype=np.float64)
        self.nbase = len(self.refcoords)
        self.vdim = 1
        self.base = self

    # ---- entity dof tables (1-based) -------------------------------------------------
    @property
    def vertexdof_indices(self):
        return tuple((i,) for i in range(1, self.refshape.nvertices + 1))

    @property
    def edgedof_interior_indices(self):
        return _INTERIOR[(self.shape, self.order)][0]

    @property
    def facedof_interior_indices(self):
        return _INTERIOR[(self.shape, self.order)][1]

    @property
    def volumedof_interior_indices(self):
        return _INTERIOR[(self.shape, self.order)][2]

    @property
    def edgedof_indices(self):
        v, e = self.vertexdof_indices, self.edgedof_interior_indices
        return tuple(tuple(sum((v[vn - 1] for vn in edge), ())) + tuple(e[k])
                     for k, edge in enumerate(self.refshape.edges))

    @property
    def facedof_indices(self):
        v, e, f = self.vertexdof_indices, self.edgedof_interior_indices, self.facedof_interior_indices
        out = []
        for k, face in enumerate(self.refshape.faces):
            d = []
            for vn in face:
                d.extend(v[vn - 1])
            for en in self.refshape.face_edgenrs[k]:
                d.extend(e[en - 1])
            d.extend(f[k])
            out.append(tuple(d))
        return tuple(out)

    @property
    def facetdof_indices(self):
        """dirichlet_facetdof_indices, src/interpolations.jl:480-482"""
        if self.rdim == 3:
            return self.facedof_indices
        if self.rdim == 2:
            return self.edgedof_indices
        return self.vertexdof_indices

    def boundarydof_indices(self, kind):
        return {"facet": self.facetdof_indices, "face": self.facedof_indices,
                "edge": self.edgedof_indices, "vertex": self.vertexdof_indices}[kind]

    # ---- values / gradients ------------------------------------------------------------
    def value_and_gradient(self, xi):
        """xi: (rdim,) -> N (nbase,), dN (nbase, rdim)."""
        xi = np.asarray(xi, dtype=np.float64)
        n, rdim = self.nbase, self.rdim
        N = np.zeros(n)
        dN = np.zeros((n, rdim))
        s, o = self.shape, self.order
        if s in ("line", "quadrilateral", "hexahedron"):
            for i in range(n):
                f = [_l1d(o, self.refcoords[i, d], xi[d]) for d in range(rdim)]
                N[i] = np.prod([fv for fv, _ in f])
                for d in range(rdim):
                    g = f[d][1]
                    for e in range(rdim):
                        if e != d:
                            g = g * f[e][0]
                    dN[i, d] = g
        elif s == "triangle":
            x, y = xi
            g = 1 - x - y
            if o == 1:
                N[:] = [x, y, g]
                dN[:] = [[1, 0], [0, 1], [-1, -1]]
            else:
                N[:] = [x * (2 * x - 1), y * (2 * y - 1), g * (2 * g - 1), 4 * x * y, 4 * y * g, 4 * x * g]
                dN[:] = [[4 * x - 1, 0], [0, 4 * y - 1], [-(4 * g - 1), -(4 * g - 1)],
                         [4 * y, 4 * x], [-4 * y, 4 * g - 4 * y], [4 * g - 4 * x, -4 * x]]
        elif s == "tetrahedron":
            x, y, z = xi
            g = 1 - x - y - z
            if o == 1:
                N[:] = [g, x, y, z]
                dN[:] = [[-1, -1, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1]]
            else:
                N[:] = [(2 * g - 1) * g, x * (2 * x - 1), y * (2 * y - 1), z * (2 * z - 1),
                        4 * x * g, 4 * x * y, 4 * y * g, 4 * z * g, 4 * x * z, 4 * y * z]
                a = -(4 * g - 1)
                dN[:] = [[a, a, a], [4 * x - 1, 0, 0], [0, 4 * y - 1, 0], [0, 0, 4 * z - 1],
                         [4 * g - 4 * x, -4 * x, -4 * x], [4 * y, 4 * x, 0],
                         [-4 * y, 4 * g - 4 * y, -4 * y], [-4 * z, -4 * z, 4 * g - 4 * z],
                         [4 * z, 0, 4 * x], [0, 4 * z, 4 * y]]
        else:
            raise ValueError(s)
        return N, dN

    def __pow__(self, vdim):
        return VectorLagrange(self, int(vdim))

    def __repr__(self):
        return f"Lagrange({self.shape},{self.order})"


class VectorLagrange:
    """VectorizedInterpolation `ip^vdim` (src/interpolations.jl:1769-1825)."""

    def __init__(self, base, vdim):
        self.base = base
        self.vdim = vdim
        self.shape = base.shape
        self.order = base.order
        self.rdim = base.rdim
        self.refshape = base.refshape
        self.nbase = base.nbase * vdim

    def __repr__(self):
        return f"{self.base!r}^{self.vdim}"


def geometric_interpolation(shape):
    """Linear Lagrange of the cell (grid cells in scope are all first-order)."""
    return Lagrange(shape, 1)
