"""Pin the CPU oracle to the reference's own golden values (CPU only, no GPU).

Every expected value below is a literal taken from the reference's tests/docs
(path:line under the reference tree is cited per test).  The oracle may only be
trusted for parity checks of the CUDA path because these pass.
"""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

import oracle as O


def test_dofs_line_golden():
    # test/test_dofs.jl:60-92
    nodes = np.array([[0.0, 0.0], [1.0, 1.0], [2.0, 0.0]])
    grid = O.Grid("line", [(1, 2), (2, 3)], nodes)
    dh = O.DofHandler(grid).add("x", O.Lagrange("line", 1) ** 2).close()
    assert list(dh.celldofs(1)) == [1, 2, 3, 4]
    assert list(dh.celldofs(2)) == [3, 4, 5, 6]
    dh = O.DofHandler(grid).add("x", O.Lagrange("line", 2) ** 2).close()
    assert list(dh.celldofs(1)) == [1, 2, 3, 4, 5, 6]
    assert list(dh.celldofs(2)) == [3, 4, 7, 8, 9, 10]
    dh = O.DofHandler(grid).add("u", O.Lagrange("line", 2) ** 3).add("th", O.Lagrange("line", 2) ** 3).close()
    assert list(dh.celldofs(1)) == list(range(1, 19))
    assert list(dh.celldofs(2)) == [4, 5, 6, 19, 20, 21, 22, 23, 24, 13, 14, 15, 25, 26, 27, 28, 29, 30]


def test_dofs_shell_quads_golden():
    # test/test_dofs.jl:95-126
    nodes = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [2, 0, 0], [2, 2, 0]], dtype=float)
    grid = O.Grid("quadrilateral", [(1, 2, 3, 4), (2, 5, 6, 3)], nodes)
    q1, q2 = O.Lagrange("quadrilateral", 1), O.Lagrange("quadrilateral", 2)
    dh = O.DofHandler(grid).add("u", q1 ** 3).add("th", q1 ** 3).close()
    assert list(dh.celldofs(1)) == list(range(1, 25))
    assert list(dh.celldofs(2)) == [4, 5, 6, 25, 26, 27, 28, 29, 30, 7, 8, 9,
                                    16, 17, 18, 31, 32, 33, 34, 35, 36, 19, 20, 21]
    dh = O.DofHandler(grid).add("u", q2).add("th", q2).close()
    assert list(dh.celldofs(1)) == list(range(1, 19))
    assert list(dh.celldofs(2)) == [2, 19, 20, 3, 21, 22, 23, 6, 24, 11, 25, 26, 12, 27, 28, 29, 15, 30]


def test_dof_range_golden():
    # test/test_dofs.jl:40-48
    grid = O.generate_grid("triangle", (10, 10))
    dh = O.DofHandler(grid).add("u", O.Lagrange("triangle", 2) ** 2).add("p", O.Lagrange("triangle", 1)).close()
    assert dh.dof_range("u") == (1, 12)
    assert dh.dof_range("p") == (13, 15)


def test_dofs_and_dirichlet_two_fields_golden():
    # test/test_dofs.jl:235-257 (without the AffineConstraint on dof 13, out of scope)
    grid = O.generate_grid("quadrilateral", (2, 1))
    q1 = O.Lagrange("quadrilateral", 1)
    dh = O.DofHandler(grid).add("v", q1 ** 2).add("s", q1).close()
    assert list(dh.celldofs(1)) == list(range(1, 13))
    assert list(dh.celldofs(2)) == [3, 4, 13, 14, 15, 16, 5, 6, 10, 17, 18, 11]
    ch = O.ConstraintHandler(dh)
    ch.add(O.Dirichlet("v", grid.facetsets["left"], lambda x, t: 0, [2]))
    ch.add(O.Dirichlet("s", grid.facetsets["left"], lambda x, t: 0))
    ch.close()
    assert list(ch.prescribed_dofs) == [2, 8, 9, 12]       # reference: [2, 8, 9, 12, 13] with the affine dof 13


def test_node_bc_golden():
    # test/test_constraints.jl:84-101
    grid = O.generate_grid("triangle", (1, 1))
    nodeset = [i + 1 for i, x in enumerate(grid.nodes) if x[1] == -1 or x[0] == -1]
    p1 = O.Lagrange("triangle", 1)
    dh = O.DofHandler(grid).add("u", p1 ** 2).add("p", p1).close()
    ch = O.ConstraintHandler(dh)
    ch.add(O.Dirichlet("u", nodeset, lambda x, t: x, [1, 2], kind="node"))
    ch.add(O.Dirichlet("p", nodeset, lambda x, t: 0, 1, kind="node"))
    ch.close()
    assert list(ch.prescribed_dofs) == list(range(1, 10))
    assert list(ch.inhomogeneities) == [-1, -1, 1, -1, -1, 1, 0, 0, 0]


def test_edge_bc_hex_golden():
    # test/test_constraints.jl:155-172: edge set {x1 == -1 and x3 == -1} of a 1-cell hex = local edge 4 (v4,v1)
    grid = O.generate_grid("hexahedron", (1, 1, 1))
    h1 = O.Lagrange("hexahedron", 1)
    dh = O.DofHandler(grid).add("u", h1 ** 3).add("p", h1).close()
    ch = O.ConstraintHandler(dh)
    ch.add(O.Dirichlet("u", [(1, 4)], lambda x, t: x, [1, 2, 3], kind="edge"))
    ch.close()
    assert list(ch.prescribed_dofs) == [1, 2, 3, 10, 11, 12]
    assert list(ch.inhomogeneities) == [-1.0, -1.0, -1.0, -1.0, 1.0, -1.0]


def test_edge_bc_shell_golden():
    # test/test_constraints.jl:175-200: bottom edges (x2 == 0) are local edge 1 of both quads
    nodes = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [2, 0, 0], [2, 2, 0]], dtype=float)
    grid = O.Grid("quadrilateral", [(1, 2, 3, 4), (2, 5, 6, 3)], nodes)
    q2 = O.Lagrange("quadrilateral", 2)
    dh = O.DofHandler(grid).add("u", q2).add("th", q2).close()
    ch = O.ConstraintHandler(dh)
    ch._skip_update = True
    dbc = O.Dirichlet("th", [(1, 1), (2, 1)], lambda x, t: (0.0,), [1], kind="edge")
    ch.add(dbc)
    # geometry is embedded (sdim 3, rdim 2): only the dof set is pinned by the reference here
    pre = sorted(ch._pre)
    assert pre == [10, 11, 14, 25, 27]


def test_assemble_apply_kat_golden():
    # test/test_assembler_extensions.jl:44-86: Line (2,), Q1, u(left) = 1
    grid = O.generate_grid("line", (2,))
    dh = O.DofHandler(grid).add("u", O.Lagrange("line", 1)).close()
    K = O.allocate_matrix(dh)
    # pattern of the reference: I = [1,1,2,2,2,3,3], J = [1,2,1,2,3,2,3] (CSR listing); as CSC:
    assert list(K.colptr) == [1, 3, 6, 8]
    assert list(K.rowval) == [1, 2, 1, 2, 3, 2, 3]
    f = np.zeros(3)
    ch = O.ConstraintHandler(dh)
    ch.add(O.Dirichlet("u", grid.facetsets["left"], lambda x, t: 1))
    ch.close()
    O.start_assemble(K, f)
    ke = np.array([[-1.0, 1.0], [2.0, -1.0]])
    fe = np.array([1.0, 2.0])
    O.assemble_cell(K, f, [1, 2], ke, fe)
    O.assemble_cell(K, f, [3, 2], ke, fe)
    dense = K.toscipy().toarray()
    assert np.allclose(dense, [[-1, 1, 0], [2, -2, 2], [0, 1, -1]])
    assert np.allclose(f, [1.0, 4.0, 1.0])
    ch.apply(K, f)
    dense = K.toscipy().toarray()
    assert np.allclose(dense, [[4 / 3, 0, 0], [0, -2, 2], [0, 1, -1]], rtol=1e-15)
    assert np.allclose(f, [4 / 3, 2.0, 1.0], rtol=1e-15)


def test_scatter_zero_skip_and_missing_entry():
    # test/test_assemble.jl:171-214 semantics
    grid = O.generate_grid("line", (2,))
    dh = O.DofHandler(grid).add("u", O.Lagrange("line", 1)).close()
    K = O.allocate_matrix(dh)
    ke = np.zeros((2, 2))
    O.assemble_cell(K, None, [1, 3], ke)              # (1,3) is not in the pattern but the values are zero: fine
    ke[0, 1] = 1.0
    try:
        O.assemble_cell(K, None, [1, 3], ke)
        raise AssertionError("expected MissingEntryError")
    except O.MissingEntryError:
        pass


def test_matrix_norm_golden_p2_triangles():
    # docs/src/topics/assembly.md:111-115,340-356: norm(K.nzval) == 1138.8803468514259
    grid = O.generate_grid("triangle", (100, 100))
    ip = O.Lagrange("triangle", 2)
    dh = O.DofHandler(grid).add("u", ip).close()
    assert dh.ndofs == 40401
    K = O.allocate_matrix(dh)
    cv = O.CellValues(O.QuadratureRule("triangle", 2), ip)
    O.assemble_global(dh, cv, K, None, "heat")
    ref = 1138.8803468514259
    assert abs(np.linalg.norm(K.nzval) - ref) / ref < 1e-13


def test_heat_tutorial_golden():
    # docs/src/literate-tutorials/heat_equation.jl:59-114,181-234: norm(u) == 3.307743912641305
    grid = O.generate_grid("quadrilateral", (20, 20))
    ip = O.Lagrange("quadrilateral", 1)
    dh = O.DofHandler(grid).add("u", ip).close()
    K = O.allocate_matrix(dh)
    f = np.zeros(dh.ndofs)
    cv = O.CellValues(O.QuadratureRule("quadrilateral", 2), ip)
    ch = O.ConstraintHandler(dh)
    boundary = np.concatenate([grid.facetsets[k] for k in ("left", "right", "top", "bottom")])
    ch.add(O.Dirichlet("u", boundary, lambda x, t: 0))
    ch.close()
    O.assemble_global(dh, cv, K, f, "heat")
    ch.apply(K, f)
    u = spla.spsolve(K.toscipy().tocsc(), f)
    ref = 3.307743912641305
    assert abs(np.linalg.norm(u) - ref) / ref < 1e-12


def test_nnz_formulas():
    # SURVEY A4 / src/Dofs/sparsity_pattern.jl:1136-1204: Q1 nnz = vdim^2 prod(3 n_i + 1), Q2 nnz = vdim^2 (8n+1)^3
    g = O.generate_grid("hexahedron", (3, 4, 5))
    dh = O.DofHandler(g).add("u", O.Lagrange("hexahedron", 1)).close()
    assert O.allocate_matrix(dh).nnz == 10 * 13 * 16
    g = O.generate_grid("hexahedron", (2, 2, 2))
    dh = O.DofHandler(g).add("u", O.Lagrange("hexahedron", 2) ** 3).close()
    assert dh.ndofs == 3 * 5 ** 3
    assert O.allocate_matrix(dh).nnz == 9 * 17 ** 3


def test_quadrature_and_partition_of_unity():
    # test/test_quadrules.jl:22-100 (weights sum to the reference volume), test/test_interpolations.jl (sum N = 1)
    vol = {"line": 2.0, "quadrilateral": 4.0, "hexahedron": 8.0, "triangle": 0.5, "tetrahedron": 1 / 6}
    for shape, v in vol.items():
        for order in (1, 2, 3):
            qr = O.QuadratureRule(shape, order)
            assert abs(qr.weights.sum() - v) < 1e-12
            for o in (1, 2):
                ip = O.Lagrange(shape, o)
                for xi in qr.points:
                    N, dN = ip.value_and_gradient(xi)
                    assert abs(N.sum() - 1) < 1e-14
                    assert np.abs(dN.sum(axis=0)).max() < 1e-13
                # Kronecker property at the reference coordinates
                V = np.array([ip.value_and_gradient(x)[0] for x in ip.refcoords])
                assert np.allclose(V, np.eye(ip.nbase), atol=1e-14)
                # gradient vs central finite differences
                x0 = qr.points[0]
                _, dN = ip.value_and_gradient(x0)
                for d in range(ip.rdim):
                    e = np.zeros(ip.rdim)
                    e[d] = 1e-6
                    fd = (ip.value_and_gradient(x0 + e)[0] - ip.value_and_gradient(x0 - e)[0]) / 2e-6
                    assert np.allclose(fd, dN[:, d], atol=1e-8)


def test_neohooke_tangent_is_derivative_of_residual():
    # hyperelasticity.jl:241-276: ke = d ge / d ue (consistency of the closed-form dS/dC with the AD tangent)
    grid = O.generate_grid("tetrahedron", (1, 1, 1), (0, 0, 0), (1, 1, 1))
    ip = O.Lagrange("tetrahedron", 2) ** 3
    cv = O.CellValues(O.QuadratureRule("tetrahedron", 4), ip)
    lam, mu = O.lame(10.0, 0.3)
    p = {"lambda": lam, "mu": mu, "b": (0.0, -0.5, 0.0)}
    rng = np.random.default_rng(0)
    x = grid.nodes[grid.cells[:1] - 1]
    u = 0.05 * rng.standard_normal((1, 30))
    ke, ge = O.element_neohooke(cv, x, p, u)
    h = 1e-6
    for j in range(0, 30, 7):
        du = np.zeros_like(u)
        du[0, j] = h
        gp = O.element_neohooke(cv, x, p, u + du)[1]
        gm = O.element_neohooke(cv, x, p, u - du)[1]
        assert np.allclose((gp - gm)[0] / (2 * h), ke[0, :, j], rtol=1e-6, atol=1e-8)
    assert np.allclose(ke, np.swapaxes(ke, 1, 2), atol=1e-12)


def test_c_port_matches_numpy_oracle():
    # the C restatement (bench cpu_baseline) against the numpy oracle, heat + elasticity, 1 and 4 threads
    from oracle import cport
    nel = (6, 5, 4)
    for vdim, elem, params in [(1, "heat", {"k": 1.5, "source": 0.5}),
                               (3, "elasticity", dict(zip(("lambda", "mu"), O.lame(200e9, 0.3)), b=(0.0, 0.0, -1.0)))]:
        grid = O.perturb_grid(O.generate_grid("hexahedron", nel), nel, (-1,) * 3, (1,) * 3, 0.2)
        ip = O.Lagrange("hexahedron", 1)
        ip = ip ** vdim if vdim > 1 else ip
        dh = O.DofHandler(grid).add("u", ip).close()
        cv = O.CellValues(O.QuadratureRule("hexahedron", 2), ip)
        K1, f1 = O.allocate_matrix(dh), np.zeros(dh.ndofs)
        O.assemble_global(dh, cv, K1, f1, elem, params)
        for nt in (1, 4):
            K2, f2 = O.allocate_matrix(dh), np.zeros(dh.ndofs)
            cport.assemble(dh, cv, K2, f2, elem, params, nthreads=nt)
            scale = np.abs(K1.nzval).max()
            assert np.all(np.abs(K2.nzval - K1.nzval) <= 1e-13 * scale)
            assert np.all(np.abs(f2 - f1) <= 1e-13 * max(np.abs(f1).max(), 1e-300))


def test_c_port_neohooke_matches_numpy_oracle():
    # Neo-Hooke tangent + residual of the C restatement (bench cpu_baseline for the tetrahedral configuration)
    from oracle import cport
    lam, mu = O.lame(10.0, 0.3)
    params = {"lambda": lam, "mu": mu, "b": (0.0, -0.5, 0.1)}
    for shape, nel, order, qo in (("tetrahedron", (3, 3, 2), 2, 4), ("hexahedron", (3, 2, 2), 1, 2)):
        grid = O.perturb_grid(O.generate_grid(shape, nel), nel, (-1,) * 3, (1,) * 3, 0.2)
        ip = O.Lagrange(shape, order) ** 3
        dh = O.DofHandler(grid).add("u", ip).close()
        cv = O.CellValues(O.QuadratureRule(shape, qo), ip)
        u = 0.02 * np.sin(0.7 * np.arange(dh.ndofs))
        K1, f1 = O.allocate_matrix(dh), np.zeros(dh.ndofs)
        O.assemble_global(dh, cv, K1, f1, "neohooke", params=params, u=u)
        K2, f2 = O.allocate_matrix(dh), np.zeros(dh.ndofs)
        cport.assemble(dh, cv, K2, f2, "neohooke", params, nthreads=3, u=u)
        assert np.all(np.abs(K2.nzval - K1.nzval) <= 1e-13 * np.abs(K1.nzval).max())
        assert np.all(np.abs(f2 - f1) <= 1e-13 * np.abs(f1).max())


def _hyperelasticity_problem(O_):
    """docs/src/literate-tutorials/hyperelasticity.jl:329-391: grid, values, Dirichlet data, Neumann set"""
    N, L = 10, 1.0
    grid = O_.generate_grid("tetrahedron", (N, N, N), (0.0, 0.0, 0.0), (L, L, L))
    ip = O_.Lagrange("tetrahedron", 1) ** 3
    dh = O_.DofHandler(grid).add("u", ip).close()
    cv = O_.CellValues(O_.QuadratureRule("tetrahedron", 1), ip)
    fv = O_.FacetValues(O_.FacetQuadratureRule("tetrahedron", 1), ip)
    th = np.pi / 3

    def rotation(x, t):
        return t * np.array([0.0, L / 2 - x[1] + (x[1] - L / 2) * np.cos(th) - (x[2] - L / 2) * np.sin(th),
                             L / 2 - x[2] + (x[1] - L / 2) * np.sin(th) + (x[2] - L / 2) * np.cos(th)])

    ch = O_.ConstraintHandler(dh)
    ch.add(O_.Dirichlet("u", grid.facetsets["right"], lambda x, t: [0.0, 0.0, 0.0], [1, 2, 3]))
    ch.add(O_.Dirichlet("u", grid.facetsets["left"], rotation, [1, 2, 3]))
    ch.close()
    ch.update(0.5)
    gamma_n = np.concatenate([grid.facetsets[k] for k in ("top", "bottom", "front", "back")])
    E, nu = 10.0, 0.3
    mp = {"mu": E / (2 * (1 + nu)), "lambda": E * nu / ((1 + nu) * (1 - 2 * nu)), "b": (0.0, -0.5, 0.0)}
    return grid, dh, cv, fv, ch, gamma_n, mp


def test_hyperelasticity_tutorial_golden():
    # docs/src/literate-tutorials/hyperelasticity.jl:241-291 (element + traction), :329-442 (Newton loop):
    # norm(u) == 4.761404305083876.  Pins the Neo-Hooke tangent + residual, the facet traction term,
    # inhomogeneous Dirichlet data and apply_zero!.  The tutorial solves with CG; a direct solve reaches the same
    # Newton fixed point (tolerance 1e-8 on the residual).
    grid, dh, cv, fv, ch, gamma_n, mp = _hyperelasticity_problem(O)
    K = O.allocate_matrix(dh)
    g = np.zeros(dh.ndofs)
    un = np.zeros(dh.ndofs)
    ch.apply_vec(un)
    du = np.zeros(dh.ndofs)
    for it in range(31):
        u = un + du
        O.assemble_global(dh, cv, K, g, "neohooke", params=mp, u=u)
        O.assemble_facets(dh, fv, g, gamma_n, "normal_traction", -0.1)
        ch.apply_zero(K, g)
        if np.linalg.norm(g) < 1e-8:
            break
        ddu = spla.spsolve(K.toscipy().tocsc(), g)
        ch.apply_vec(ddu, applyzero=True)
        du -= ddu
    else:
        raise AssertionError("Newton did not converge")
    ref = 4.761404305083876
    assert abs(np.linalg.norm(u) - ref) / ref < 1e-7, np.linalg.norm(u)


def test_facet_values_areas_and_normals():
    # sum of dGamma over a boundary set = its area; normals point outward (src/FEValues/FacetValues.jl:128-154)
    for shape, nel, area in (("hexahedron", (3, 2, 2), 4.0), ("tetrahedron", (2, 2, 2), 4.0),
                             ("quadrilateral", (3, 2), 2.0), ("triangle", (3, 2), 2.0)):
        grid = O.generate_grid(shape, nel)
        for order in (1, 2):
            ip = O.Lagrange(shape, order)
            fv = O.FacetValues(O.FacetQuadratureRule(shape, 2), ip)
            expect = {"left": -1.0, "right": 1.0}
            for name, sign in expect.items():
                pairs = np.asarray(grid.facetsets[name]).reshape(-1, 2)
                tot, nsum = 0.0, 0.0
                for facet in np.unique(pairs[:, 1]):
                    cells = pairs[pairs[:, 1] == facet, 0] - 1
                    n, dG = O.reinit_facet(fv, grid.nodes[grid.cells[cells] - 1], int(facet))
                    tot += dG.sum()
                    assert np.allclose(n[..., 0], sign) and np.allclose(n[..., 1:], 0.0)
                assert abs(tot - area) < 1e-13, (shape, name, tot)


def _renumber_testdh():
    # test/test_dofs.jl:230-246 (testdhch) without the AffineConstraint on dof 13 (out of scope)
    grid = O.generate_grid("quadrilateral", (2, 1))
    q1 = O.Lagrange("quadrilateral", 1)
    dh = O.DofHandler(grid).add("v", q1 ** 2).add("s", q1).close()
    ch = O.ConstraintHandler(dh)
    ch.add(O.Dirichlet("v", grid.facetsets["left"], lambda x, t: 0, [2]))
    ch.add(O.Dirichlet("s", grid.facetsets["left"], lambda x, t: 0))
    ch.close()
    return dh, ch


RENUMBER_GOLDENS = [
    # (order, target blocks, celldofs(1), celldofs(2), prescribed dofs without the image of the affine dof 13)
    # test/test_dofs.jl:260-316
    ("fieldwise", None, [1, 2, 3, 4, 5, 6, 7, 8, 13, 14, 15, 16], [3, 4, 9, 10, 11, 12, 5, 6, 14, 17, 18, 15], [2, 8, 13, 16]),
    ("fieldwise", [2, 1], [7, 8, 9, 10, 11, 12, 13, 14, 1, 2, 3, 4], [9, 10, 15, 16, 17, 18, 11, 12, 2, 5, 6, 3], [1, 4, 8, 14]),
    ("componentwise", None, [1, 7, 2, 8, 3, 9, 4, 10, 13, 14, 15, 16], [2, 8, 5, 11, 6, 12, 3, 9, 14, 17, 18, 15], [7, 10, 13, 16]),
    ("componentwise", [3, 1, 2], [13, 1, 14, 2, 15, 3, 16, 4, 7, 8, 9, 10], [14, 2, 17, 5, 18, 6, 15, 3, 8, 11, 12, 9], [1, 4, 7, 10]),
]


def test_renumber_goldens():
    # renumber!(dh, ch, DofOrder.FieldWise / ComponentWise): literal celldofs and prescribed dofs of test/test_dofs.jl:260-316
    for order, tb, c1, c2, pre in RENUMBER_GOLDENS:
        dh, ch = _renumber_testdh()
        O.renumber(dh, O.renumber_permutation(dh, order, tb), ch)
        assert list(dh.celldofs(1)) == c1 and list(dh.celldofs(2)) == c2, order
        assert list(ch.prescribed_dofs) == pre, order
    # roundtrip with an arbitrary permutation and its inverse (test/test_dofs.jl:183-206)
    dh, ch = _renumber_testdh()
    cd0, pre0 = dh.cell_dofs.copy(), np.array(ch.prescribed_dofs)
    perm = np.random.default_rng(0).permutation(dh.ndofs) + 1
    iperm = np.empty_like(perm)
    iperm[perm - 1] = np.arange(1, dh.ndofs + 1)
    O.renumber(dh, perm, ch)
    O.renumber(dh, iperm, ch)
    assert np.array_equal(dh.cell_dofs, cd0) and np.array_equal(ch.prescribed_dofs, pre0)


def test_local_application_of_bc_golden():
    # test/test_constraints.jl:1425-1610 "local application of bc": 5x5 quads, Q1, u(left) = 0, u(right) = 1, element with
    # conductivity k = cellid and source b = 1/cellid; literal goldens norm(u) = 3.8249286998373586 (apply!) and
    # 0.06401424182205259 (apply_zero!), and the identities between apply!, apply_assemble! and apply_local! + assemble!
    grid = O.generate_grid("quadrilateral", (5, 5))
    ip = O.Lagrange("quadrilateral", 1)
    dh = O.DofHandler(grid).add("u", ip).close()
    ch = O.ConstraintHandler(dh)
    ch.add(O.Dirichlet("u", grid.facetsets["left"], lambda x, t: 0))
    ch.add(O.Dirichlet("u", grid.facetsets["right"], lambda x, t: 1))
    ch.close()
    ch.update(0.0)
    cv = O.CellValues(O.QuadratureRule("quadrilateral", 2), ip)
    Kes, fes = O.element_matrices(dh, cv, "heat", {"k": 1.0, "source": 1.0})
    ids = np.arange(1, grid.ncells + 1, dtype=float)
    Kes *= ids[:, None, None]
    fes /= ids[:, None]
    pd = ch.prescribed_dofs - 1
    fd = np.setdiff1d(np.arange(dh.ndofs), pd)
    for azero, golden in ((False, 3.8249286998373586), (True, 0.06401424182205259)):
        Ks, Kc, Kl = O.allocate_matrix(dh), O.allocate_matrix(dh), O.allocate_matrix(dh)
        fs, fc, fl = np.zeros(dh.ndofs), np.zeros(dh.ndofs), np.zeros(dh.ndofs)
        for c in range(grid.ncells):
            dofs = dh.cell_dofs[c]
            O.assemble_cell(Ks, fs, dofs, Kes[c], fes[c])
            O.apply_assemble(Kc, fc, ch, dofs, Kes[c].copy(), fes[c].copy(), applyzero=azero)
            ke, fe = Kes[c].copy(), fes[c].copy()
            O.apply_local(ke, fe, dofs, ch, applyzero=azero)
            O.assemble_cell(Kl, fl, dofs, ke, fe)
        ch.apply(Ks, fs, applyzero=azero)
        As, Ac, Al = Ks.toscipy().toarray(), Kc.toscipy().toarray(), Kl.toscipy().toarray()
        assert np.array_equal(Ac, Al) and np.array_equal(fc, fl)
        assert np.allclose(As[np.ix_(fd, fd)], Ac[np.ix_(fd, fd)], rtol=1e-13, atol=0)
        assert np.allclose(fs[fd], fc[fd], rtol=1e-12, atol=1e-14)
        for A, f in ((As, fs), (Ac, fc)):
            P = A[np.ix_(pd, pd)]
            assert np.array_equal(P, np.diag(np.diag(P)))
            assert np.all(A[np.ix_(pd, fd)] == 0) and np.all(A[np.ix_(fd, pd)] == 0)
            assert np.allclose(f[pd] / np.diag(P), 0.0 if azero else ch.inhomogeneities, rtol=1e-14, atol=0)
            u = np.linalg.solve(A, f)
            assert abs(np.linalg.norm(u) - golden) < 1e-12 * golden, (azero, np.linalg.norm(u))


def _poisson_errors_oracle(shape, order, N):
    # test/integration/convergence_test_utils.jl: -lap u = f with u = prod cos(pi x_i / 2) on [-1, 1]^d, u prescribed on
    # every facet set; the right-hand side is the nodal interpolant of f integrated with the mass matrix (the kernel menu
    # has no x-dependent source), which keeps the orders p + 1 (L2) and p (H1)
    dim = {"quadrilateral": 2, "hexahedron": 3}[shape]
    grid = O.generate_grid(shape, (N,) * dim)
    ip = O.Lagrange(shape, order)
    dh = O.DofHandler(grid).add("u", ip).close()
    cv = O.CellValues(O.QuadratureRule(shape, max(2 * order - 1, 2)), ip)
    ana = lambda x: np.prod(np.cos(np.pi * np.asarray(x) / 2), axis=-1)     # noqa: E731
    ch = O.ConstraintHandler(dh)
    boundary = np.concatenate([grid.facetsets[k] for k in sorted(grid.facetsets)])
    ch.add(O.Dirichlet("u", boundary, lambda x, t: ana(x)))
    ch.close()
    ch.update(0.0)
    # dof coordinates: interpolate the coordinate field (exact for the multilinear geometry)
    geo = O.Lagrange(shape, 1)
    xd = np.zeros((dh.ndofs, dim))
    for a in range(ip.nbase):
        M, _ = geo.value_and_gradient(ip.refcoords[a])
        xd[dh.cell_dofs[:, a] - 1] = np.einsum("j,cjd->cd", M, grid.nodes[grid.cells - 1])
    K, Mm = O.allocate_matrix(dh), O.allocate_matrix(dh)
    O.assemble_global(dh, cv, K, np.zeros(dh.ndofs), "heat", {"k": 1.0, "source": 0.0})
    O.assemble_global(dh, cv, Mm, None, "mass", {"rho": 1.0})
    f = Mm.toscipy() @ (dim * np.pi ** 2 / 4 * ana(xd))
    ch.apply(K, f)
    u = spla.spsolve(K.toscipy().tocsc(), f)
    ch.apply_vec(u)
    dNdx, dO = O.reinit(cv, grid.nodes[grid.cells - 1])
    ue = u[dh.cell_dofs - 1]
    xq = np.einsum("qa,cad->cqd", cv.N, xd[dh.cell_dofs - 1])
    uh = np.einsum("qa,ca->cq", cv.N, ue)
    gh = np.einsum("ca,cqad->cqd", ue, dNdx)
    ua = ana(xq)
    ga = np.stack([-np.pi / 2 * np.tan(np.pi * xq[..., d] / 2) * ua for d in range(dim)], axis=-1)
    return (np.sqrt(np.sum((ua - uh) ** 2 * dO)), np.sqrt(np.sum(np.sum((ga - gh) ** 2, axis=-1) * dO)), np.abs(ua - uh).max())


def test_poisson_convergence_rates():
    # test/integration/convergence_test_utils.jl:176-207: L2 rate = order + 1, H1 rate = order (atol 0.1), pointwise bounds
    for shape, order, N in (("quadrilateral", 1, 21), ("quadrilateral", 2, 7), ("hexahedron", 1, 11)):
        l1, h1, m1 = _poisson_errors_oracle(shape, order, N)
        l2, h2, m2 = _poisson_errors_oracle(shape, order, 2 * N)
        assert m1 < 3e-2 and m2 < 1e-2      # the reference's 1e-2 / 5e-3 hold for the exact source; the interpolated one costs a factor 3
        assert abs(np.log(l1 / l2) / np.log(2) - (order + 1)) < 0.1, (shape, order, np.log(l1 / l2) / np.log(2))
        assert abs(np.log(h1 / h2) / np.log(2) - order) < 0.1, (shape, order, np.log(h1 / h2) / np.log(2))


def test_apply_rhs_golden():
    # test/test_apply_rhs.jl:5-92: apply!(K, f, ch) and get_rhs_data + apply!(A, ch) + apply_rhs!(data, g, ch) give the same solve
    grid = O.generate_grid("quadrilateral", (20, 20))
    ip = O.Lagrange("quadrilateral", 1)
    dh = O.DofHandler(grid).add("u", ip).close()
    cv = O.CellValues(O.QuadratureRule("quadrilateral", 2), ip)
    ch = O.ConstraintHandler(dh)
    ch.add(O.Dirichlet("u", np.concatenate([grid.facetsets["left"], grid.facetsets["right"]]), lambda x, t: 0))
    ch.add(O.Dirichlet("u", np.concatenate([grid.facetsets["top"], grid.facetsets["bottom"]]), lambda x, t: 2))
    ch.close()
    ch.update(0.0)
    K, A = O.allocate_matrix(dh), O.allocate_matrix(dh)
    f, g = np.zeros(dh.ndofs), np.zeros(dh.ndofs)
    O.assemble_global(dh, cv, K, f, "heat")
    O.assemble_global(dh, cv, A, g, "heat")
    data = O.get_rhs_data(ch, A)
    ch.apply(K, f)
    ch.apply(A)
    O.apply_rhs(data, g, ch)
    assert np.array_equal(K.nzval, A.nzval)
    assert np.allclose(f, g, rtol=1e-15, atol=1e-15)
    u1 = spla.spsolve(K.toscipy().tocsc(), f)
    u2 = spla.spsolve(A.toscipy().tocsc(), g)
    assert np.allclose(u1, u2, rtol=1e-14, atol=0)
    assert abs(u1[ch.prescribed_dofs - 1] - ch.inhomogeneities).max() < 1e-14


def test_affine_constraints_golden():
    # Groundwork for the next round (oracle only; the CUDA library has no affine constraints yet).
    # test/test_constraints.jl:323-412: in-place apply! with condensation == the explicitly condensed system C'KC a_f = C'(f - Kg)
    grid = O.generate_grid("line", (10,))
    dh = O.DofHandler(grid).add("u", O.Lagrange("line", 1)).close()
    n = dh.ndofs
    # nonsymmetric matrix condensation (:330-341)
    ch = O.AffineConstraintHandler(dh)
    ch.add(O.AffineConstraint(1, [(3, 2.0)], 0.0))
    ch.add(O.AffineConstraint(2, [(4, 3.0)], 0.0))
    ch.close()
    C, _ = ch.create_constraint_matrix()
    K = O.dense_pattern(n)
    K.nzval[:] = np.arange(1.0, n * n + 1)                 # reshape(1.0:(n^2), n, n), column-major like nzval
    Kd = K.toscipy().toarray()
    ch.apply(K)
    fd = ch.free_dofs - 1
    Cd = C.toarray()
    assert np.array_equal(K.toscipy().toarray()[np.ix_(fd, fd)], Cd.T @ Kd @ Cd)
    test_acs = [
        [O.AffineConstraint(4, [(7, 1.0)], 0.0)],
        [O.AffineConstraint(2, [(5, 1.0), (6, 2.0)], 1.0)],
        [O.AffineConstraint(2, [(9, 1.0)], 0.0), O.AffineConstraint(3, [(9, 1.0)], 0.0)],
        [O.AffineConstraint(2, [(7, 3.0), (8, 1.0)], -1.0), O.AffineConstraint(4, [(9, -1.0)], 2.0)],
    ]
    for acs in test_acs:
        ch = O.AffineConstraintHandler(dh)
        ch.add(O.Dirichlet("u", grid.facetsets["left"], lambda x, t: 0.0))
        for ac in acs:
            ch.add(ac)
        ch.close()
        ch.update(0.0)
        C, g = ch.create_constraint_matrix()
        C = C.toarray()
        K = O.allocate_matrix_condensed(dh, ch)
        f = np.zeros(n)
        f[-1] = 1.0
        for c in range(grid.ncells):
            O.assemble_cell(K, None, dh.cell_dofs[c], 2.0 * np.array([[1.0, -1.0], [-1.0, 1.0]]))
        Kd = K.toscipy().toarray()
        aa = C @ np.linalg.solve((C.T @ Kd @ C), C.T @ (f - Kd @ g)) + g
        ch.apply(K, f)
        a = ch.apply_vec(np.linalg.solve(K.toscipy().toarray(), f))
        assert np.allclose(a, aa, rtol=1e-12, atol=1e-13)
        for ac in acs:                                         # the constraint holds in the solution
            assert abs(a[ac.constrained_dof - 1] - (ac.b + sum(v * a[d - 1] for d, v in ac.entries))) < 1e-12
    # error paths (:1417-1422)
    ch = O.AffineConstraintHandler(dh)
    ch.add(O.AffineConstraint(1, [(2, 1.0)], 0.0))
    ch.add(O.AffineConstraint(2, [(3, 1.0)], 0.0))
    try:
        ch.close()
        assert False, "nested constraints must be rejected"
    except ValueError as e:
        assert "nested affine constraints currently not supported" in str(e)


def test_condensed_pattern_equals_pattern_of_extended_cells():
    # the pattern of allocate_matrix(dh, ch) is the pattern of allocate_matrix(dh) plus, per cell,
    # ext x ext with ext = (cell dofs that are not affinely constrained) + (masters of the cell's affinely constrained dofs)
    # -- what lets the incidence-based device builder produce it from one extra padded dof list per cell
    rng = np.random.default_rng(7)
    for shape, nel, order in (("quadrilateral", (5, 4), 1), ("triangle", (4, 4), 2), ("hexahedron", (3, 2, 2), 1)):
        grid = O.generate_grid(shape, nel)
        dh = O.DofHandler(grid).add("u", O.Lagrange(shape, order)).close()
        n = dh.ndofs
        slaves = rng.choice(np.arange(1, n + 1), size=max(3, n // 6), replace=False)
        pool = np.setdiff1d(np.arange(1, n + 1), slaves)
        ch = O.AffineConstraintHandler(dh)
        for s in slaves:
            masters = rng.choice(pool, size=rng.integers(0, 4), replace=False)
            ch.add(O.AffineConstraint(int(s), [(int(m), float(rng.random() + 0.5)) for m in masters], float(rng.random())))
        ch.close()
        K = O.allocate_matrix_condensed(dh, ch)
        K0 = O.allocate_matrix(dh)
        pairs = set(zip(K0.rowval.tolist(), np.repeat(np.arange(1, n + 1), np.diff(K0.colptr)).tolist()))
        for c in range(grid.ncells):
            ext = []
            for d in dh.cell_dofs[c]:
                co = ch._coeffs(d)
                ext += [int(d)] if co is None else [m for m, _ in co]
            pairs.update((r, cc) for r in ext for cc in ext)
        cols = np.repeat(np.arange(1, n + 1), np.diff(K.colptr))
        assert pairs == set(zip(K.rowval.tolist(), cols.tolist())), shape


def test_c_port_coloured_scheme_matches_numpy_oracle():
    # the coloured threading scheme of docs/src/literate-howto/threaded_assembly.jl:330-377 in the C port (CPU baseline):
    # valid colourings for generate_grid meshes, same K and f as the sequential oracle, bitwise reproducible run to run
    from oracle import cport
    lam, mu = O.lame(200e9, 0.3)
    cases = [("hexahedron", (7, 6, 5), 1, 1, 2, "heat", {"k": 2.0, "source": 3.0}),
             ("hexahedron", (4, 3, 3), 2, 3, 3, "elasticity", {"lambda": lam, "mu": mu, "b": (0.0, 0.0, -1.0)}),
             ("tetrahedron", (3, 3, 2), 2, 1, 2, "heat", None), ("quadrilateral", (6, 5), 1, 1, 2, "heat", None),
             ("triangle", (5, 4), 2, 1, 2, "heat", None)]
    for shape, nel, order, vdim, qo, el, params in cases:
        dim = len(nel)
        grid = O.perturb_grid(O.generate_grid(shape, nel), nel, (-1,) * dim, (1,) * dim, 0.2)
        ip = O.Lagrange(shape, order)
        ip = ip ** vdim if vdim > 1 else ip
        dh = O.DofHandler(grid).add("u", ip).close()
        cv = O.CellValues(O.QuadratureRule(shape, qo), ip)
        colors = cport.structured_coloring(grid, nel)
        assert colors[0] == {"hexahedron": 8, "tetrahedron": 48, "quadrilateral": 4, "triangle": 8}[shape]
        K0, K1, K2 = O.allocate_matrix(dh), O.allocate_matrix(dh), O.allocate_matrix(dh)
        f0, f1, f2 = np.zeros(dh.ndofs), np.zeros(dh.ndofs), np.zeros(dh.ndofs)
        O.assemble_global(dh, cv, K0, f0, el, params)
        cport.assemble(dh, cv, K1, f1, el, params, nthreads=4, colors=colors)
        cport.assemble(dh, cv, K2, f2, el, params, nthreads=2, colors=colors)
        assert np.abs(K1.nzval - K0.nzval).max() <= 1e-13 * np.abs(K0.nzval).max()
        assert np.abs(f1 - f0).max() <= 1e-13 * max(np.abs(f0).max(), 1e-300)
        assert np.array_equal(K1.nzval, K2.nzval) and np.array_equal(f1, f2)      # threaded_assembly.jl:397-410


def _cook_problem(n=50):
    """the quadratic / linear problem of incompressible_elasticity.jl:386-446 with the oracle: (K, f, ch, dh)"""
    g = O.cook_grid(n, n)
    ipu, ipp = O.Lagrange("triangle", 2) ** 2, O.Lagrange("triangle", 1)
    dh = O.DofHandler(g).add("u", ipu).add("p", ipp).close()
    qr = O.QuadratureRule("triangle", 3)
    cvu, cvp = O.CellValues(qr, ipu), O.CellValues(qr, ipp)
    nu_ = 0.5
    G = 1.0 / (2 * (1 + nu_))
    K = O.allocate_matrix(dh)
    O.assemble_up(dh, cvu, cvp, K, G, 0.0)              # K = Emod nu / (3 (1 - 2 nu)) = Inf: 1 / K = 0
    f = np.zeros(dh.ndofs)
    fv = O.FacetValues(O.FacetQuadratureRule("triangle", 3), ipu)
    nuu = ipu.nbase
    pairs = g.facetsets["traction"]
    for facet in np.unique(pairs[:, 1]):
        cells = pairs[pairs[:, 1] == facet, 0] - 1
        fe = O.facet_element(fv, g.nodes[g.cells[cells] - 1], int(facet), "traction", (0.0, 1.0 / 16.0))
        np.add.at(f, dh.cell_dofs[cells][:, :nuu] - 1, fe)
    ch = O.ConstraintHandler(dh)
    ch.add(O.Dirichlet("u", g.facetsets["clamped"], lambda x, t: [0.0, 0.0], [1, 2]))
    ch.close()
    ch.update(0.0)
    return g, dh, K, f, ch, (cvu, cvp, G)


def test_incompressible_elasticity_tutorial_golden():
    # docs/src/literate-tutorials/incompressible_elasticity.jl:477: norm(u2) = 919.1284143115702 -- pins the two-field dof
    # numbering, MultiFieldCellValues (both fields on one rule), the mixed u-p element, the traction facet term and apply!
    import scipy.sparse.linalg as spla
    g, dh, K, f, ch, _ = _cook_problem(50)
    ch.apply(K, f)
    u = spla.spsolve(K.toscipy().tocsc(), f)
    assert abs(np.linalg.norm(u) - 919.1284143115702) <= 1e-8 * 919.1284143115702


@pytest.mark.parametrize("nel,vdim", [((5, 4, 3), 1), ((3, 4, 5), 3), ((1, 1, 1), 1), ((7, 2, 6), 2)])
def test_c_setup_matches_numpy_oracle(nel, vdim):
    # oracle/cpu_setup.c (the set-up of the full-size CPU reference arm) against the numpy oracle: nodes, cells, dofs, pattern
    from oracle import cport
    g, dh, K = cport.hex_q1_problem(nel, vdim)
    og = O.perturb_grid(O.generate_grid("hexahedron", nel), nel, (-1.0,) * 3, (1.0,) * 3, 0.2)
    ip = O.Lagrange("hexahedron", 1)
    ip = ip ** vdim if vdim > 1 else ip
    odh = O.DofHandler(og).add("u", ip).close()
    oK = O.allocate_matrix(odh)
    assert np.array_equal(g.cells, og.cells) and np.allclose(g.nodes, og.nodes, rtol=0, atol=1e-15)
    assert np.array_equal(dh.cell_dofs, odh.cell_dofs) and dh.ndofs == odh.ndofs
    assert np.array_equal(K.colptr, oK.colptr) and np.array_equal(K.rowval, oK.rowval)
