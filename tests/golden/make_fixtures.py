"""Generate tests/golden/fixtures.npz: input/output vectors of small seeded cases, computed by the CPU oracle.

The reference is a Julia package and cannot run in the build container or on the GPU box (no Julia toolchain), so
these vectors come from `oracle/` -- which tests/test_oracle_goldens.py pins to the reference's own literal goldens
(`reference_literals.json` lists them with file:line).  The fixtures let the GPU parity tests check the CUDA path
against committed numbers without executing the oracle, and guard the oracle itself against regressions.

    python tests/golden/make_fixtures.py        # rewrites fixtures.npz next to this file
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

# name -> (shape, nel, order, vdim, qr_order, element, params)
CASES = {
    "hex_q1_heat": ("hexahedron", (4, 3, 2), 1, 1, 2, "heat", {"k": 1.5, "source": 0.7}),
    "hex_q2v3_elasticity": ("hexahedron", (2, 2, 2), 2, 3, 3, "elasticity", {"E": 200e9, "nu": 0.3, "b": (0.0, 0.0, -1.0)}),
    "tet_p2v3_neohooke": ("tetrahedron", (2, 2, 2), 2, 3, 4, "neohooke", {"E": 10.0, "nu": 0.3, "b": (0.0, -0.5, 0.0)}),
    "quad_q2_heat_dirichlet": ("quadrilateral", (5, 4), 2, 1, 3, "heat", {"k": 1.0, "source": 1.0}),
    "tri_p2_mass": ("triangle", (4, 3), 2, 1, 3, "mass", {"rho": 2.0}),
}


def state(grid, dh, vdim):
    """deterministic smooth displacement evaluated at the vertex dofs (others zero)"""
    u = np.zeros(dh.ndofs)
    nv = grid.cells.shape[1]
    for c in range(grid.ncells):
        for a in range(nv):
            x = grid.nodes[grid.cells[c, a] - 1]
            for k in range(vdim):
                u[dh.cell_dofs[c, a * vdim + k] - 1] = 0.03 * np.sin(1.3 * x[(k + 1) % len(x)] + 0.4 * k)
    return u


def compute(name):
    shape, nel, order, vdim, qo, element, p = CASES[name]
    dim = len(nel)
    left, right = (-1.0,) * dim, (1.0,) * dim
    grid = O.perturb_grid(O.generate_grid(shape, nel, left, right), nel, left, right, 0.2)
    ip = O.Lagrange(shape, order)
    ip = ip ** vdim if vdim > 1 else ip
    dh = O.DofHandler(grid).add("u", ip).close()
    K = O.allocate_matrix(dh)
    f = np.zeros(dh.ndofs)
    cv = O.CellValues(O.QuadratureRule(shape, qo), ip)
    params, u = dict(p), None
    if element in ("elasticity", "neohooke"):
        lam, mu = O.lame(p["E"], p["nu"])
        params = {"lambda": lam, "mu": mu, "b": p["b"]}
    if element == "neohooke":
        u = state(grid, dh, vdim)
    O.assemble_global(dh, cv, K, f, element, params=params, u=u)
    out = {"nodes": grid.nodes, "cells": grid.cells, "cell_dofs": dh.cell_dofs, "colptr": K.colptr, "rowval": K.rowval,
           "nzval": K.nzval.copy(), "f": f.copy()}
    if u is not None:
        out["u"] = u
    if name.endswith("dirichlet"):
        ch = O.ConstraintHandler(dh)
        boundary = np.concatenate([np.asarray(grid.facetsets[k]).reshape(-1, 2) for k in ("left", "top")])
        ch.add(O.Dirichlet("u", boundary, lambda x, t: 1.0 + 0.5 * x[0]))
        ch.close()
        ch.apply(K, f)
        out.update(prescribed=np.asarray(ch.prescribed_dofs), inhom=np.asarray(ch.inhomogeneities),
                   nzval_applied=K.nzval.copy(), f_applied=f.copy())
    return out


def main():
    data = {}
    for name in CASES:
        for k, v in compute(name).items():
            data[f"{name}/{k}"] = np.asarray(v)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures.npz")
    np.savez_compressed(path, **data)
    print(path, os.path.getsize(path), "bytes", len(data), "arrays")


if __name__ == "__main__":
    main()
