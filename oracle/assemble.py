"""Global assembly loop + CSC scatter (oracle; test infrastructure only).

Restates the reference's `assemble_global!` loop (docs/src/literate-tutorials/
heat_equation.jl:181-204): `start_assemble` zero-fill (src/assembler.jl:287-291,
src/arrayutils.jl:118-152), CellCache gather (src/iterators.jl:72-87) and the
scatter semantics of `assemble!` (src/assembler.jl:322-331,347-457):
K[dofs[i],dofs[j]] += Ke[i,j], f[dofs[i]] += fe[i]; exact zeros are skipped; a
non-zero aimed at an entry missing from the pattern is an error (:459-467);
duplicate dofs in one cell add twice.  Contributions are summed in ascending
cell order (`np.add.at` is sequential), which is the reference's serial order.
"""
import numpy as np

from .element import ELEMENTS

__all__ = ["start_assemble", "assemble_cell", "assemble_global", "element_matrices", "ea_mul", "MissingEntryError"]


class MissingEntryError(KeyError):
    """src/assembler.jl:459-467 `_missing_sparsity_pattern_error`"""


def start_assemble(K, f=None, fillzero=True):
    if fillzero:
        K.nzval[:] = 0.0
        if f is not None:
            f[:] = 0.0
    return K, f


def assemble_cell(K, f, dofs, Ke, fe=None):
    """One `assemble!(assembler, dofs, Ke, fe)`; dofs 1-based."""
    dofs = np.asarray(dofs, dtype=np.int64)
    Ke = np.asarray(Ke, dtype=np.float64)
    n = len(dofs)
    rows = np.repeat(dofs[:, None], n, axis=1)
    cols = np.repeat(dofs[None, :], n, axis=0)
    pos = K.lookup(rows.ravel(), cols.ravel())
    v = Ke.ravel()
    nz = v != 0
    if np.any((pos < 0) & nz):
        k = np.argwhere((pos < 0) & nz)[0, 0]
        raise MissingEntryError(f"entry ({rows.ravel()[k]}, {cols.ravel()[k]}) is missing in the sparsity pattern")
    np.add.at(K.nzval, pos[nz], v[nz])
    if fe is not None and f is not None:
        np.add.at(f, dofs - 1, np.asarray(fe, dtype=np.float64))


def assemble_global(dh, cv, K, f, element="heat", params=None, u=None, chunk=4096, fillzero=True, cells=None):
    """Whole `assemble_global!` loop.  `cells`: optional 0-based subset (partitioned assembly)."""
    fn = ELEMENTS[element] if isinstance(element, str) else element
    start_assemble(K, f, fillzero)
    grid = dh.grid
    n = dh.ndofs_per_cell
    assert n == cv.nbase, "oracle: element must cover all dofs of the cell"
    ids = np.arange(grid.ncells) if cells is None else np.asarray(cells)
    for s in range(0, len(ids), chunk):
        sl = ids[s:s + chunk]
        x = grid.nodes[grid.cells[sl] - 1]
        cd = dh.cell_dofs[sl]
        ue = None if u is None else u[cd - 1]
        Ke, fe = fn(cv, x, params, ue)
        rows = np.repeat(cd[:, :, None], n, axis=2)
        cols = np.repeat(cd[:, None, :], n, axis=1)
        pos = K.lookup(rows.ravel(), cols.ravel())
        v = Ke.ravel()
        nz = v != 0
        if np.any((pos < 0) & nz):
            raise MissingEntryError("an entry is missing in the sparsity pattern")
        np.add.at(K.nzval, pos[nz], v[nz])
        if f is not None:
            np.add.at(f, (cd - 1).ravel(), fe.ravel())
    return K, f


def element_matrices(dh, cv, element="heat", params=None, u=None, chunk=4096):
    """Kes (ncells, n, n) and fes (ncells, n): the element routine of every cell, nothing scattered -- the "element
    assembly" storage of docs/src/literate-howto/gpu_assembly.jl:265-285."""
    fn = ELEMENTS[element] if isinstance(element, str) else element
    grid = dh.grid
    n = dh.ndofs_per_cell
    Kes = np.zeros((grid.ncells, n, n))
    fes = np.zeros((grid.ncells, n))
    for s in range(0, grid.ncells, chunk):
        sl = slice(s, min(s + chunk, grid.ncells))
        x = grid.nodes[grid.cells[sl] - 1]
        ue = None if u is None else u[dh.cell_dofs[sl] - 1]
        Kes[sl], fes[sl] = fn(cv, x, params, ue)
    return Kes, fes


def ea_mul(dh, Kes, x):
    """y = sum_e P_e' Ke P_e x in ascending cell order (gpu_assembly.jl:287-304)"""
    y = np.zeros(dh.ndofs)
    for c in range(dh.grid.ncells):
        d = dh.cell_dofs[c] - 1
        np.add.at(y, d, Kes[c] @ x[d])
    return y
