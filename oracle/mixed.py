"""Mixed u-p (incompressible elasticity) element and Cook's membrane (oracle; test infrastructure only).

Restates docs/src/literate-tutorials/incompressible_elasticity.jl: create_cook_grid (:160-172), MultiFieldCellValues(qr, (u, p))
(:178-190; src/FEValues/CellValues.jl:229-298: all fields on one quadrature rule and geometric mapping), assemble_up!
(:266-311) and solve (:386-425).  Pinned on the tutorial's own literal norm(u2) = 919.1284143115702 (:477, quadratic / linear
triangles, 50 x 50, nu = 0.5) in tests/test_oracle_goldens.py.
"""
import numpy as np

from .element import reinit
from .grid import Grid, _generate_nodes, generate_grid

__all__ = ["cook_grid", "element_up", "assemble_up"]


def cook_grid(nx, ny):
    """generate_grid(Triangle, (nx, ny), corners) with the four corners of Cook's membrane + the "clamped" / "traction" facet
    sets of addfacetset!(grid, name, x -> x[1] ~ 0 / 48) (all vertices of the facet satisfy the predicate)"""
    g0 = generate_grid("triangle", (nx, ny))
    corners = np.array([[0.0, 0.0], [48.0, 44.0], [48.0, 60.0], [0.0, 44.0]])
    nodes = _generate_nodes("quadrilateral", (nx + 1, ny + 1), corners)
    g = Grid("triangle", g0.cells, nodes, dict(g0.facetsets))
    facets = [(1, 2), (2, 3), (3, 1)]
    for name, x0 in (("clamped", 0.0), ("traction", 48.0)):
        out = []
        for ci in range(g.ncells):
            for fi, (a, b) in enumerate(facets):
                if all(abs(nodes[g.cells[ci, v - 1] - 1, 0] - x0) <= 1e-8 * max(1.0, abs(x0)) for v in (a, b)):
                    out.append((ci + 1, fi + 1))
        g.facetsets[name] = np.array(sorted(out), dtype=np.int64).reshape(-1, 2)
    return g


def element_up(cvu, cvp, x, G, invK):
    """Ke (ncells, n, n), local dofs = [u dofs (a, c) -> a * dim + c | p dofs]: K_uu = int 2G dev3d(sym grad phi_I) : dev3d(sym
    grad phi_J), K_pu = -int psi_i div phi_J, K_pp = -int psi_i psi_j / K (incompressible_elasticity.jl:266-293)"""
    dNdx, dOm = reinit(cvu, x)                      # (cells, nq, nbu, dim), (cells, nq)
    nc, nq, nbu, dim = dNdx.shape
    nbp = cvp.N.shape[1]
    nu = nbu * dim
    Ke = np.zeros((nc, nu + nbp, nu + nbp))
    gg = np.einsum("cqae,cqbe,cq->cab", dNdx, dNdx, dOm)
    t1 = np.einsum("cqae,cqbd,cq->caebd", dNdx, dNdx, dOm)          # int g_a[e] g_b[d]
    K5 = 0.5 * np.swapaxes(t1, 2, 4) - t1 / 3.0                       # g_a[d] g_b[e] / 2 - g_a[e] g_b[d] / 3
    for e in range(dim):
        K5[:, :, e, :, e] += 0.5 * gg
    Ke[:, :nu, :nu] = 2.0 * G * K5.reshape(nc, nu, nu)
    Kpu = -np.einsum("qi,cqbd,cq->cibd", cvp.N, dNdx, dOm).reshape(nc, nbp, nu)
    Ke[:, nu:, :nu] = Kpu
    Ke[:, :nu, nu:] = np.swapaxes(Kpu, 1, 2)
    Ke[:, nu:, nu:] = -invK * np.einsum("qi,qj,cq->cij", cvp.N, cvp.N, dOm)
    return Ke


def assemble_up(dh, cvu, cvp, K, G, invK):
    from .assemble import assemble_cell
    g = dh.grid
    x = g.nodes[g.cells - 1]
    Ke = element_up(cvu, cvp, x, G, invK)
    for c in range(g.ncells):
        assemble_cell(K, None, dh.cell_dofs[c], Ke[c])
    return K
