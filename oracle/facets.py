"""FacetQuadratureRule, FacetValues and the Neumann / traction facet loop (oracle; test infrastructure only).

Restates, for the Lagrange menu of this repo:
  * `FacetQuadratureRule{shape}(order)`        src/Quadrature/quadrature.jl:205-238
    (`create_facet_quad_rule`                   src/FEValues/facet_integrals.jl:39-47)
  * `facet_to_element_transformation`          src/FEValues/facet_integrals.jl:102-110,139-146,168-178,209-217
  * `weighted_normal`                          src/FEValues/facet_integrals.jl:122-131,158-166,192-203,230-239
  * `reinit!(fv::FacetValues, cell, x, facet)` src/FEValues/FacetValues.jl:128-154
    (detJ = |weighted normal| > 0, n = unit normal, dGamma = detJ * w)
  * the traction loop of the hyperelasticity tutorial, docs/src/literate-tutorials/hyperelasticity.jl:278-291
    (`ge[i] -= (du_i . t) dGamma`, t = tn * n) and `assemble!(f, dofs, fe)` src/assembler.jl:338-345.
"""
import numpy as np

from .interpolations import geometric_interpolation
from .quadrature import QuadratureRule

__all__ = ["FacetQuadratureRule", "FacetValues", "reinit_facet", "facet_element", "assemble_facets"]

_FACET_SHAPE = {"quadrilateral": "line", "triangle": "line", "hexahedron": "quadrilateral", "tetrahedron": "triangle"}
_NFACETS = {"quadrilateral": 4, "triangle": 3, "hexahedron": 6, "tetrahedron": 4}


def facet_to_element(shape, facet, p):
    """facet: 1-based local facet number; p: point of the facet's reference shape."""
    if shape == "quadrilateral":
        x = p[0]
        return {1: (x, -1.0), 2: (1.0, x), 3: (-x, 1.0), 4: (-1.0, -x)}[facet]
    if shape == "triangle":
        x = (p[0] + 1.0) / 2
        return {1: (1.0 - x, x), 2: (0.0, 1.0 - x), 3: (x, 0.0)}[facet]
    if shape == "hexahedron":
        x, y = p
        return {1: (y, x, -1.0), 2: (x, -1.0, y), 3: (1.0, x, y), 4: (-x, 1.0, y), 5: (-1.0, y, x), 6: (x, y, 1.0)}[facet]
    if shape == "tetrahedron":
        x, y = p
        return {1: (1.0 - x - y, y, 0.0), 2: (y, 0.0, 1.0 - x - y), 3: (x, y, 1.0 - x - y), 4: (0.0, 1.0 - x - y, y)}[facet]
    raise ValueError(shape)


def weighted_normal(shape, facet, J):
    """J: (..., sdim, rdim) -> (..., sdim)."""
    c = lambda k: J[..., :, k - 1]
    if shape == "quadrilateral":
        if facet == 1: return np.stack([J[..., 1, 0], -J[..., 0, 0]], -1)
        if facet == 2: return np.stack([J[..., 1, 1], -J[..., 0, 1]], -1)
        if facet == 3: return np.stack([-J[..., 1, 0], J[..., 0, 0]], -1)
        if facet == 4: return np.stack([-J[..., 1, 1], J[..., 0, 1]], -1)
    if shape == "triangle":
        if facet == 1: return np.stack([-(J[..., 1, 0] - J[..., 1, 1]), J[..., 0, 0] - J[..., 0, 1]], -1)
        if facet == 2: return np.stack([-J[..., 1, 1], J[..., 0, 1]], -1)
        if facet == 3: return np.stack([J[..., 1, 0], -J[..., 0, 0]], -1)
    if shape == "hexahedron":
        pairs = {1: (2, 1), 2: (1, 3), 3: (2, 3), 4: (3, 1), 5: (3, 2), 6: (1, 2)}
        a, b = pairs[facet]
        return np.cross(c(a), c(b))
    if shape == "tetrahedron":
        if facet == 1: return np.cross(c(2), c(1))
        if facet == 2: return np.cross(c(1), c(3))
        if facet == 3: return np.cross(c(1) - c(3), c(2) - c(3))
        if facet == 4: return np.cross(c(3), c(2))
    raise ValueError((shape, facet))


class FacetQuadratureRule:
    """One QuadratureRule per local facet, points in the CELL's reference coordinates."""

    def __init__(self, shape, order):
        self.shape = shape
        base = QuadratureRule(_FACET_SHAPE[shape], order)
        w = base.weights / 2 if shape == "triangle" else base.weights   # quadrature.jl:231-232
        self.rules = []
        for facet in range(1, _NFACETS[shape] + 1):
            pts = np.array([facet_to_element(shape, facet, p) for p in base.points], dtype=np.float64)
            self.rules.append(QuadratureRule(shape, weights=w.copy(), points=pts))

    @property
    def nfacets(self):
        return len(self.rules)


class FacetValues:
    def __init__(self, fqr, ip, ip_geo=None):
        self.fqr = fqr
        self.ip = ip
        self.base = ip.base
        self.vdim = ip.vdim
        self.ip_geo = (ip_geo.base if ip_geo is not None else geometric_interpolation(ip.shape))
        nf, nq = fqr.nfacets, fqr.rules[0].nq
        self.N = np.zeros((nf, nq, self.base.nbase))
        self.dNdxi = np.zeros((nf, nq, self.base.nbase, self.base.rdim))
        self.M = np.zeros((nf, nq, self.ip_geo.nbase))
        self.dMdxi = np.zeros((nf, nq, self.ip_geo.nbase, self.ip_geo.rdim))
        self.w = np.zeros((nf, nq))
        for f, rule in enumerate(fqr.rules):
            self.w[f] = rule.weights
            for q in range(nq):
                self.N[f, q], self.dNdxi[f, q] = self.base.value_and_gradient(rule.points[q])
                self.M[f, q], self.dMdxi[f, q] = self.ip_geo.value_and_gradient(rule.points[q])

    @property
    def nbase(self):
        return self.base.nbase * self.vdim


def reinit_facet(fv, x, facet):
    """x: (n, ngeo, sdim) coordinates of n cells, all on local facet `facet` (1-based).
    Returns unit normals (n, nq, sdim) and dGamma (n, nq)."""
    shape = fv.fqr.shape
    J = np.einsum("cja,qjb->cqab", x, fv.dMdxi[facet - 1])
    wn = weighted_normal(shape, facet, J)
    det = np.sqrt(np.sum(wn * wn, axis=-1))
    if not np.all(det > 0):
        raise ArithmeticError("det(J) is not positive on a facet")
    return wn / det[..., None], det * fv.w[facet - 1][None, :]


def facet_element(fv, x, facet, kind, params):
    """fe (n, nbase) of the facet integral.
    kind 'normal_traction': fe[(i,c)] = p * int N_i n_c dGamma      (params: scalar p; tutorial: p = -tn)
    kind 'traction':        fe[(i,c)] = int N_i t_c dGamma          (params: vector t of length vdim)
    kind 'flux':            fe[i]     = q * int N_i dGamma          (scalar field, params: scalar q)"""
    n, dG = reinit_facet(fv, x, facet)
    N = fv.N[facet - 1]                                  # (nq, nb)
    if kind == "flux":
        assert fv.vdim == 1
        return float(params) * np.einsum("qi,cq->ci", N, dG)
    if kind == "normal_traction":
        t = float(params) * n                            # (n, nq, sdim)
    elif kind == "traction":
        t = np.broadcast_to(np.asarray(params, dtype=np.float64), n.shape)
    else:
        raise ValueError(kind)
    assert fv.vdim == t.shape[-1]
    fe = np.einsum("qi,cqk,cq->cik", N, t, dG)
    return fe.reshape(fe.shape[0], -1)


def assemble_facets(dh, fv, f, facetset, kind, params):
    """f[celldofs] += fe for every (cell, facet) of `facetset` (1-based pairs): `assemble!(f, dofs, fe)`."""
    pairs = np.asarray(facetset, dtype=np.int64).reshape(-1, 2)
    grid = dh.grid
    for facet in np.unique(pairs[:, 1]):
        cells = pairs[pairs[:, 1] == facet, 0] - 1
        x = grid.nodes[grid.cells[cells] - 1]
        fe = facet_element(fv, x, int(facet), kind, params)
        np.add.at(f, dh.cell_dofs[cells] - 1, fe)
    return f
