"""Committed golden vectors (tests/golden/): the oracle must reproduce them (CPU), and the CUDA path is checked
against them without executing the oracle (GPU).  `tests/golden/make_fixtures.py` regenerates fixtures.npz."""
import json
import os

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIX = np.load(os.path.join(HERE, "fixtures.npz"))
LIT = json.load(open(os.path.join(HERE, "reference_literals.json")))
NAMES = sorted({k.split("/")[0] for k in FIX.files})
RTOL = 1e-12


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.abs(b).max() if b.size else 0.0
    return bool(np.all(np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)) + rtol * scale))


def test_fixture_file_is_complete():
    assert len(NAMES) == 5
    for n in NAMES:
        for k in ("nodes", "cells", "cell_dofs", "colptr", "rowval", "nzval", "f"):
            assert f"{n}/{k}" in FIX.files
    for key in ("norm_K_nzval_p2_triangles_100x100", "heat_tutorial_norm_u", "hyperelasticity_tutorial_norm_u"):
        assert "source" in LIT[key]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_fixtures(name):
    import sys
    sys.path.insert(0, HERE)
    import make_fixtures
    out = make_fixtures.compute(name)
    for k, v in out.items():
        ref = FIX[f"{name}/{k}"]
        if np.issubdtype(ref.dtype, np.integer):
            assert np.array_equal(np.asarray(v), ref), k       # numbering / pattern: bit-exact
        else:
            assert close(v, ref, 1e-13), k


def test_literals_match_the_oracle_tests():
    # the JSON is the single list of reference literals; the oracle tests use the same numbers
    src = open(os.path.join(os.path.dirname(HERE), "test_oracle_goldens.py")).read()
    for key in ("norm_K_nzval_p2_triangles_100x100", "heat_tutorial_norm_u", "hyperelasticity_tutorial_norm_u"):
        assert repr(LIT[key]["value"]) in src


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_path_matches_fixtures(name):
    import sys
    sys.path.insert(0, HERE)
    import make_fixtures
    import ferrite_b200 as fb
    shape, nel, order, vdim, qo, element, p = make_fixtures.CASES[name]
    ct = {"hexahedron": fb.Hexahedron, "tetrahedron": fb.Tetrahedron, "quadrilateral": fb.Quadrilateral, "triangle": fb.Triangle}[shape]
    ctx = fb.default_context(0)
    g = fb.Grid.from_arrays(ct, FIX[f"{name}/cells"], FIX[f"{name}/nodes"])      # the fixture's (perturbed) grid as arrays
    ip = fb.Lagrange(ct, order) ** vdim
    dh = fb.close_(fb.add_(fb.DofHandler(g), "u", ip))
    assert np.array_equal(dh.cell_dofs, FIX[f"{name}/cell_dofs"])
    K = fb.allocate_matrix(dh)
    assert np.array_equal(K.colptr, FIX[f"{name}/colptr"]) and np.array_equal(K.rowval, FIX[f"{name}/rowval"])
    f = ctx.zeros(dh.ndofs)
    cv = fb.CellValues(fb.QuadratureRule(ct, qo), ip)
    if element == "heat":
        elem = fb.HeatElement(p["k"], p["source"])
    elif element == "mass":
        elem = fb.MassElement(p["rho"])
    else:
        E, nu = p["E"], p["nu"]
        lam, mu = E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
        elem = (fb.ElasticityElement if element == "elasticity" else fb.NeoHookeElement)(lam=lam, mu=mu, b=p["b"])
    u = None
    if f"{name}/u" in FIX.files:
        import torch
        u = torch.from_numpy(FIX[f"{name}/u"]).to(f.device)
    fb.assemble_(fb.start_assemble(K, f), elem, cv, u=u)
    ctx.synchronize()
    assert close(K.nzval.cpu().numpy(), FIX[f"{name}/nzval"])
    assert close(f.cpu().numpy(), FIX[f"{name}/f"])
    if f"{name}/prescribed" in FIX.files:
        ch = fb.ConstraintHandler.from_arrays(dh, FIX[f"{name}/prescribed"], FIX[f"{name}/inhom"])
        fb.apply_(K, f, ch)
        ctx.synchronize()
        assert close(K.nzval.cpu().numpy(), FIX[f"{name}/nzval_applied"])
        assert close(f.cpu().numpy(), FIX[f"{name}/f_applied"])
