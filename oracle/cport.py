"""Build + call the C restatement of the reference loop (oracle/cpu_assemble.c).

TEST INFRASTRUCTURE ONLY (tests, smoke, and the cpu_baseline / --impl reference legs of bench.py).
The shared object is compiled with -march=native, so it is rebuilt whenever the host CPU differs from
the one it was built on (the GPU box is not the build container).
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "cpu_assemble.c")
_SRC2 = os.path.join(_HERE, "cpu_setup.c")
_lib = None


def _cpu_tag():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def build(force=False):
    """Compile oracle/cpu_assemble.c -> oracle/_build/liboracle_cpu.so (or a temp dir if not writable)."""
    out_dir = os.path.join(_HERE, "_build")
    so = os.path.join(out_dir, "liboracle_cpu.so")
    tag = os.path.join(out_dir, "cpu.tag")
    if (not force and os.path.exists(so) and os.path.exists(tag) and open(tag).read() == _cpu_tag()
            and os.path.getmtime(so) >= max(os.path.getmtime(_SRC), os.path.getmtime(_SRC2))):
        return so
    try:
        os.makedirs(out_dir, exist_ok=True)
        test = os.path.join(out_dir, ".w")
        open(test, "w").close()
        os.remove(test)
    except OSError:
        out_dir = tempfile.mkdtemp(prefix="oracle_cpu_")
        so = os.path.join(out_dir, "liboracle_cpu.so")
        tag = os.path.join(out_dir, "cpu.tag")
    cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-o", so, _SRC, _SRC2, "-lm"]
    subprocess.run(cmd, check=True)
    with open(tag, "w") as fh:
        fh.write(_cpu_tag())
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_assemble.restype = C.c_int64
        _lib.oracle_assemble_colored.restype = C.c_int64
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def max_threads():
    return lib().oracle_max_threads()


def structured_coloring(grid, nel):
    """Colours of a generate_grid mesh such that no two cells of a colour share a node (create_coloring(grid),
    src/Grid/coloring.jl:308-317, delivers such a colouring; any valid one serves the threaded loop): parity of the cell's
    (i, j, k) for quadrilaterals / hexahedra, times the sub-cell index for the 2 triangles / 6 tetrahedra of a cube.
    Returns (ncolors, color_ptr, color_cells) with 0-based int64 arrays; validity is checked."""
    sub = {"quadrilateral": 1, "hexahedron": 1, "triangle": 2, "tetrahedron": 6}[grid.shape]
    nel = tuple(nel)
    c = np.arange(grid.ncells, dtype=np.int64)
    cube, t = c // sub, c % sub
    par, stride, mult = np.zeros_like(c), 1, 1
    for nd in nel:
        par += ((cube // stride) % nd % 2) * mult
        stride *= nd
        mult *= 2
    color = par * sub + t
    ncolors = int(color.max()) + 1
    order = np.argsort(color, kind="stable").astype(np.int64)
    ptr = np.concatenate([[0], np.cumsum(np.bincount(color, minlength=ncolors))]).astype(np.int64)
    for k in range(ncolors):
        nodes = grid.cells[order[ptr[k]:ptr[k + 1]]].ravel()
        assert len(np.unique(nodes)) == len(nodes), "colouring is not valid for this grid"
    return ncolors, ptr, order


def assemble(dh, cv, K, f, element="heat", params=None, nthreads=0, u=None, colors=None):
    """Same contract as oracle.assemble_global, executed by the C port with `nthreads` OpenMP threads.
    colors = (ncolors, color_ptr, color_cells): the coloured loop (plain adds) instead of the atomic one."""
    grid = dh.grid
    n = dh.ndofs_per_cell
    assert n == cv.nbase
    if element == "heat":
        p = dict(k=1.0, source=1.0)
        p.update(params or {})
        eid, pv = 1, np.array([p["k"], p["source"]])
    elif element == "elasticity":
        b = list((params or {}).get("b") or (0.0,) * cv.vdim) + [0.0, 0.0, 0.0]
        eid, pv = 3, np.array([params["lambda"], params["mu"]] + b[:3])
    elif element == "neohooke":
        assert u is not None and cv.vdim == 3 and grid.sdim == 3
        b = list((params or {}).get("b") or (0.0,) * 3) + [0.0, 0.0, 0.0]
        eid, pv = 4, np.array([params["lambda"], params["mu"]] + b[:3])
    else:
        raise ValueError(element)
    uu = np.ascontiguousarray(u, dtype=np.float64) if u is not None else None
    cells = np.ascontiguousarray(grid.cells, dtype=np.int64)
    xyz = np.ascontiguousarray(grid.nodes, dtype=np.float64)
    cd = np.ascontiguousarray(dh.cell_dofs, dtype=np.int64)
    N = np.ascontiguousarray(cv.N)
    dN = np.ascontiguousarray(cv.dNdxi)
    dM = np.ascontiguousarray(cv.dMdxi)
    w = np.ascontiguousarray(cv.w)
    K.nzval[:] = 0.0
    if f is not None:
        f[:] = 0.0

    def ptr(a, t):
        return a.ctypes.data_as(C.POINTER(t))
    nco, cptr, ccells = (0, None, None) if colors is None else colors
    bad = lib().oracle_assemble_colored(
        C.c_int(eid), C.c_int(grid.sdim), C.c_int(cells.shape[1]), C.c_int(cv.base.nbase), C.c_int(cv.vdim), C.c_int(cv.nq),
        C.c_int64(grid.ncells), ptr(cells, C.c_int64), ptr(xyz, C.c_double), ptr(cd, C.c_int64),
        ptr(K.colptr, C.c_int64), ptr(K.rowval, C.c_int64), ptr(N, C.c_double), ptr(dN, C.c_double), ptr(dM, C.c_double),
        ptr(w, C.c_double), ptr(pv, C.c_double), ptr(uu, C.c_double) if uu is not None else None, ptr(K.nzval, C.c_double),
        ptr(f, C.c_double) if f is not None else None, C.c_int(nthreads), C.c_int(nco),
        ptr(np.ascontiguousarray(cptr, dtype=np.int64), C.c_int64) if nco else None,
        ptr(np.ascontiguousarray(ccells, dtype=np.int64), C.c_int64) if nco else None)
    if bad:
        raise ArithmeticError(f"det(J) is not positive in cell {bad}")
    return K, f


class _Plain:
    """attribute bag standing in for the numpy oracle's Grid / DofHandler / CSC in assemble()"""


def hex_q1_problem(nel, vdim=1, left=(-1.0, -1.0, -1.0), right=(1.0, 1.0, 1.0), amplitude=0.2):
    """(grid, dh, K) of generate_grid(Hexahedron, nel) + perturbation + close! of one Q1 field with `vdim` components +
    allocate_matrix, set up by the C restatement (oracle/cpu_setup.c): same arrays as the numpy oracle produces, in seconds
    at 200^3.  The returned objects carry exactly the attributes assemble() reads."""
    nx, ny, nz = (int(n) for n in nel)
    L = lib()
    sizes = [C.c_int64() for _ in range(4)]
    L.oracle_hex_q1_sizes(C.c_int64(nx), C.c_int64(ny), C.c_int64(nz), C.c_int(vdim), *[C.byref(s) for s in sizes])
    nn, nc, nd, nnz = (s.value for s in sizes)
    cells = np.empty((nc, 8), dtype=np.int64)
    xyz = np.empty((nn, 3), dtype=np.float64)
    cd = np.empty((nc, 8 * vdim), dtype=np.int64)
    colptr = np.empty(nd + 1, dtype=np.int64)
    rowval = np.empty(nnz, dtype=np.int64)
    lo, hi = np.asarray(left, dtype=np.float64), np.asarray(right, dtype=np.float64)

    def ptr(a, t):
        return a.ctypes.data_as(C.POINTER(t))
    L.oracle_hex_q1_setup.restype = C.c_int
    rc = L.oracle_hex_q1_setup(C.c_int64(nx), C.c_int64(ny), C.c_int64(nz), ptr(lo, C.c_double), ptr(hi, C.c_double), C.c_double(amplitude),
                               C.c_int(vdim), ptr(cells, C.c_int64), ptr(xyz, C.c_double), ptr(cd, C.c_int64), ptr(colptr, C.c_int64),
                               ptr(rowval, C.c_int64))
    if rc != 0:
        raise RuntimeError(f"oracle_hex_q1_setup failed ({rc})")
    grid, dh, K = _Plain(), _Plain(), _Plain()
    grid.cells, grid.nodes, grid.ncells, grid.nnodes, grid.sdim, grid.shape = cells, xyz, nc, nn, 3, "hexahedron"
    dh.grid, dh.cell_dofs, dh.ndofs, dh.ndofs_per_cell = grid, cd, nd, 8 * vdim
    K.n, K.nnz, K.colptr, K.rowval, K.nzval = nd, nnz, colptr, rowval, np.zeros(nnz)
    return grid, dh, K
