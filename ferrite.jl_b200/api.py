"""Host-side mirror of the Ferrite.jl API for the assembly path, on top of the C ABI.

Names follow the reference (Julia `f!` is spelled `f_` here): generate_grid, DofHandler / add_ /
close_, CellValues, QuadratureRule, Lagrange, allocate_matrix, start_assemble / assemble_,
ConstraintHandler / Dirichlet / update_ / apply_ / apply_zero_.  Everything that computes runs in
libferrite_b200.so on the GPU; torch is used only to own device memory and the CUDA stream.

Differences from the reference that a user must know:
  * the element routine is chosen from a menu (HeatElement, MassElement, ElasticityElement,
    NeoHookeElement) instead of being user code -- `assemble_(assembler, element, cv)` performs the whole
    `for cell in CellIterator(dh)` loop of the reference in one call;
  * matrices live on the device (`B200Matrix`): `K.colptr` / `K.rowval` are the reference's 1-based
    Int64 arrays, `K.nzval` a float64 CUDA tensor; `K.tocsc()` downloads a scipy matrix.
"""
import ctypes as C

import numpy as np

from . import _lib as L
from ._lib import FB2Error, DetJNotPositive, MissingPatternEntry  # noqa: F401

# reference shapes / cell types --------------------------------------------------------------------
Line, Triangle, Quadrilateral, Tetrahedron, Hexahedron = L.LINE, L.TRIANGLE, L.QUADRILATERAL, L.TETRAHEDRON, L.HEXAHEDRON
RefLine, RefTriangle, RefQuadrilateral, RefTetrahedron, RefHexahedron = Line, Triangle, Quadrilateral, Tetrahedron, Hexahedron
_RDIM = {Line: 1, Triangle: 2, Quadrilateral: 2, Tetrahedron: 3, Hexahedron: 3}


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _torch():
    import torch
    return torch


def _chain(obj, attr):
    """obj.<attr> and all of its ancestors (dh -> grid -> ctx)."""
    out = []
    p = getattr(obj, attr, None)
    while p is not None:
        out.append(p)
        p = getattr(p, "dh", None) or getattr(p, "grid", None) or getattr(p, "ctx", None)
    return out or [None]


def _destroy(obj, fn, parents=()):
    """Destroy a native handle unless one of its parents is already gone (the cyclic GC may finalise
    objects in any order at interpreter exit; the native child dereferences its parent)."""
    try:
        h = getattr(obj, "h", None)
        if not h:
            return
        for p in parents:
            if p is None or not getattr(p, "h", None):
                obj.h = None
                return
        getattr(L.lib, fn)(h)
        obj.h = None
    except Exception:
        pass


class Context:
    """One CUDA device + stream.  `Context(-1)` is host-only (grid / dof / constraint set-up logic)."""

    def __init__(self, device=0, use_torch_stream=True):
        self.h = C.c_void_p()
        L.call("fb2_ctx_create", int(device), C.byref(self.h))
        self.device = int(device)
        if device >= 0 and use_torch_stream:
            torch = _torch()
            torch.cuda.set_device(device)
            stream = torch.cuda.current_stream(device)
            L.call("fb2_ctx_set_stream", self.h, C.c_void_p(stream.cuda_stream))
            self._stream = stream

    def synchronize(self):
        L.call("fb2_ctx_synchronize", self.h)

    @property
    def launch_count(self):
        n = C.c_int64()
        L.call("fb2_ctx_launch_count", self.h, C.byref(n))
        return n.value

    def measure_fp64_peak(self):
        t = C.c_double()
        L.call("fb2_measure_fp64_peak", self.h, C.byref(t))
        return t.value

    def zeros(self, n):
        torch = _torch()
        return torch.zeros(int(n), dtype=torch.float64, device=f"cuda:{self.device}")

    def __del__(self):
        if _destroy is not None:      # module globals are already gone at interpreter shutdown
            _destroy(self, "fb2_ctx_destroy", ())


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


# ---- interpolation / quadrature descriptors ---------------------------------------------------------
class Lagrange:
    """Lagrange{refshape, order}(); `ip ** vdim` is the reference's `ip^vdim`."""

    def __init__(self, refshape, order, vdim=1):
        self.refshape, self.order, self.vdim = refshape, int(order), int(vdim)

    def __pow__(self, vdim):
        return Lagrange(self.refshape, self.order, int(vdim))


class QuadratureRule:
    """QuadratureRule{refshape}(order) with the reference's default rule for the shape."""

    def __init__(self, refshape, order):
        self.refshape, self.order = refshape, int(order)


# ---- grid ---------------------------------------------------------------------------------------------
class Grid:
    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle
        ct, nc, nn, nnpc, sdim = C.c_int(), C.c_int64(), C.c_int64(), C.c_int(), C.c_int()
        L.call("fb2_grid_info", self.h, C.byref(ct), C.byref(nc), C.byref(nn), C.byref(nnpc), C.byref(sdim))
        self.celltype, self.ncells, self.nnodes, self.nnpc, self.sdim = ct.value, nc.value, nn.value, nnpc.value, sdim.value

    @classmethod
    def from_arrays(cls, celltype, cells, nodes, ctx=None):
        """Grid(cells, nodes): cells (ncells, nnpc) 1-based node ids, nodes (nnodes, sdim)."""
        ctx = ctx or default_context()
        cells, nodes = _i64(cells), _f64(nodes)
        h = C.c_void_p()
        L.call("fb2_grid_from_host", ctx.h, celltype, cells.shape[0], nodes.shape[0], nodes.shape[1],
               _ptr(cells, C.c_int64), _ptr(nodes, C.c_double), C.byref(h))
        return cls(ctx, h)

    @property
    def cells(self):
        out = np.empty((self.ncells, self.nnpc), dtype=np.int64)
        L.call("fb2_grid_export", self.h, _ptr(out, C.c_int64), None)
        return out

    @property
    def nodes(self):
        out = np.empty((self.nnodes, self.sdim), dtype=np.float64)
        L.call("fb2_grid_export", self.h, None, _ptr(out, C.c_double))
        return out

    def set_coordinates(self, nodes):
        nodes = _f64(nodes)
        assert nodes.shape == (self.nnodes, self.sdim)
        L.call("fb2_grid_set_coordinates", self.h, _ptr(nodes, C.c_double))

    def upload_coordinates_async(self, host_tensor):
        """Stream-ordered device-only coordinate update from a (pinned) torch/numpy host buffer (nnodes, sdim)."""
        ptr = host_tensor.data_ptr() if hasattr(host_tensor, "data_ptr") else host_tensor.ctypes.data
        L.call("fb2_grid_upload_coordinates_async", self.h, C.c_void_p(ptr))

    def perturb(self, amplitude=0.2):
        L.call("fb2_grid_perturb", self.h, float(amplitude))
        return self

    def __del__(self):
        if _destroy is not None:      # module globals are already gone at interpreter shutdown
            _destroy(self, "fb2_grid_destroy", (getattr(self, "ctx", None),))


def generate_grid(celltype, nel, left=None, right=None, ctx=None, corners=None):
    """generate_grid(CellType, nel, left, right) or, for quadrilaterals / triangles, generate_grid(CellType, nel, corners) with
    the four corner points (src/Grid/grid_generators.jl:78-112,383-417: nodes = bilinear map of the unit grid onto the corners)"""
    if corners is not None:
        assert celltype in (Quadrilateral, Triangle) and left is None and right is None
        g = generate_grid(celltype, nel, ctx=ctx)
        xi = g.nodes                                    # the generated nodes on [-1, 1]^2 are the reference coordinates
        c = _f64(corners)
        M = np.stack([(1 - xi[:, 0]) * (1 - xi[:, 1]), (1 + xi[:, 0]) * (1 - xi[:, 1]), (1 + xi[:, 0]) * (1 + xi[:, 1]),
                      (1 - xi[:, 0]) * (1 + xi[:, 1])], axis=1) / 4.0
        g.set_coordinates(M @ c)
        return g
    ctx = ctx or default_context()
    nel = _i64(nel)
    lp = _ptr(_f64(left), C.c_double) if left is not None else None
    rp = _ptr(_f64(right), C.c_double) if right is not None else None
    h = C.c_void_p()
    L.call("fb2_grid_generate", ctx.h, celltype, _ptr(nel, C.c_int64), lp, rp, C.byref(h))
    return Grid(ctx, h)


# local vertex ids (1-based) of the facets of every cell type, src/Grid/grid.jl:196-259 (reference_facets)
_FACETS = {
    Line: ((1,), (2,)),
    Triangle: ((1, 2), (2, 3), (3, 1)),
    Quadrilateral: ((1, 2), (2, 3), (3, 4), (4, 1)),
    Tetrahedron: ((1, 3, 2), (1, 2, 4), (2, 3, 4), (1, 4, 3)),
    Hexahedron: ((1, 4, 3, 2), (1, 2, 6, 5), (2, 3, 7, 6), (3, 4, 8, 7), (1, 5, 8, 4), (5, 6, 7, 8)),
}


def _user_sets(grid, kind):
    if not hasattr(grid, "_sets"):
        grid._sets = {"facet": {}, "node": {}, "cell": {}}
    return grid._sets[kind]


def _check_setname(sets, name, grid=None):
    """src/Grid/utils.jl `_check_setname`: user sets AND, for facet sets, the sets generate_grid created ("left", "top", ...)"""
    taken = name in sets
    if not taken and grid is not None:
        n = C.c_int64()
        taken = L.lib.fb2_grid_facetset(grid.h, name.encode(), C.byref(n), None) == 0
    if taken:
        raise ValueError(f"There already exists a set with the name: {name}")


def _warn_emptyset(items, name):
    """src/Grid/utils.jl `_warn_emptyset`"""
    if len(items) == 0:
        import warnings
        warnings.warn(f"no entities added to the set with name {name}")


def _passes(f, coords, all_):
    vals = (bool(f(x)) for x in coords)
    return all(vals) if all_ else any(vals)


def addfacetset_(grid, name, f_or_pairs, all=True):
    """addfacetset!(grid, name, set_or_predicate; all = true) (src/Grid/utils.jl:42-60,119-134): with a predicate f(x), every
    (cell, local facet) whose vertex coordinates all (any) satisfy f, in lexicographic order; interior facets included."""
    sets = _user_sets(grid, "facet")
    _check_setname(sets, name, grid)
    if callable(f_or_pairs):
        nodes, cells, out = grid.nodes, grid.cells, []
        for ci in range(grid.ncells):
            for fi, verts in enumerate(_FACETS[grid.celltype]):
                if _passes(f_or_pairs, (nodes[cells[ci, v - 1] - 1] for v in verts), all):
                    out.append((ci + 1, fi + 1))
        pairs = np.asarray(out, dtype=np.int64).reshape(-1, 2)
    else:
        pairs = _i64(sorted(set(map(tuple, np.asarray(f_or_pairs, dtype=np.int64).reshape(-1, 2).tolist())))).reshape(-1, 2)
    _warn_emptyset(pairs, name)
    sets[name] = pairs
    return grid


def addnodeset_(grid, name, f_or_ids):
    """addnodeset!(grid, name, ids_or_predicate) (src/Grid/utils.jl:29-40,201-207)"""
    sets = _user_sets(grid, "node")
    _check_setname(sets, name)
    if callable(f_or_ids):
        ids = [i + 1 for i, x in enumerate(grid.nodes) if f_or_ids(x)]
    else:
        ids = sorted(set(int(i) for i in f_or_ids))
    sets[name] = _i64(ids)
    return grid


def addcellset_(grid, name, f_or_ids, all=True):
    """addcellset!(grid, name, ids_or_predicate; all = true) (src/Grid/utils.jl:21-27,188-200)"""
    sets = _user_sets(grid, "cell")
    _check_setname(sets, name)
    if callable(f_or_ids):
        nodes, cells = grid.nodes, grid.cells
        ids = [ci + 1 for ci in range(grid.ncells) if _passes(f_or_ids, (nodes[n - 1] for n in cells[ci]), all)]
    else:
        ids = sorted(set(int(i) for i in f_or_ids))
    sets[name] = _i64(ids)
    return grid


def getnodeset(grid, name):
    return _user_sets(grid, "node")[name]


def getcellset(grid, name):
    return _user_sets(grid, "cell")[name]


def getfacetset(grid, name):
    if name in _user_sets(grid, "facet"):
        return _user_sets(grid, "facet")[name]
    n = C.c_int64()
    L.call("fb2_grid_facetset", grid.h, name.encode(), C.byref(n), None)
    out = np.empty((n.value, 2), dtype=np.int64)
    L.call("fb2_grid_facetset", grid.h, name.encode(), C.byref(n), _ptr(out, C.c_int64))
    return out


# ---- DofHandler -----------------------------------------------------------------------------------------
class DofHandler:
    def __init__(self, grid):
        self.grid = grid
        self.field_names, self.field_ips = [], []
        self.h = None

    def _fields(self):
        arr = (L.Field * len(self.field_ips))()
        for k, ip in enumerate(self.field_ips):
            arr[k].order, arr[k].vdim = ip.order, ip.vdim
        return arr

    def _info(self):
        nd, ndpc, nf = C.c_int64(), C.c_int(), C.c_int()
        L.call("fb2_dh_info", self.h, C.byref(nd), C.byref(ndpc), C.byref(nf))
        self.ndofs, self.ndofs_per_cell = nd.value, ndpc.value

    @classmethod
    def from_arrays(cls, grid, fields, ndofs, cell_dofs):
        """Adopt the reference's own numbering: cell_dofs (ncells, ndofs_per_cell) 1-based."""
        dh = cls(grid)
        for name, ip in fields:
            add_(dh, name, ip)
        cd = _i64(cell_dofs)
        dh.h = C.c_void_p()
        L.call("fb2_dh_from_host", grid.h, len(dh.field_ips), dh._fields(), int(ndofs), cd.shape[1],
               _ptr(cd, C.c_int64), C.byref(dh.h))
        dh._info()
        return dh

    @property
    def cell_dofs(self):
        out = np.empty((self.grid.ncells, self.ndofs_per_cell), dtype=np.int64)
        L.call("fb2_dh_export", self.h, _ptr(out, C.c_int64))
        return out

    def __del__(self):
        if _destroy is not None:      # module globals are already gone at interpreter shutdown
            _destroy(self, "fb2_dh_destroy", _chain(self, "grid"))


class DofOrder:
    """DofOrder.FieldWise([target_blocks]) / DofOrder.ComponentWise([target_blocks]) (src/Dofs/DofRenumbering.jl:1-40)"""

    class FieldWise:
        kind = L.ORDER_FIELDWISE

        def __init__(self, target_blocks=None):
            self.target_blocks = None if target_blocks is None else [int(b) for b in target_blocks]

    class ComponentWise(FieldWise):
        kind = L.ORDER_COMPONENTWISE

    class Metis(FieldWise):
        """DofOrder.Ext{Metis}(): fill-reducing nested-dissection order (ext/FerriteMetis.jl:29-92)"""
        kind = L.ORDER_METIS

        def __init__(self):
            self.target_blocks = None


def renumber_(dh, *args):
    """renumber!(dh, order) / renumber!(dh, ch, order): order = a permutation vector (1-based, dof i becomes perm[i]),
    DofOrder.FieldWise(...) or DofOrder.ComponentWise(...).  Returns the permutation that was applied (1-based).
    Renumber before allocate_matrix / start_assemble: matrices of the old numbering are stale (as in the reference)."""
    ch, order = (None, args[0]) if len(args) == 1 else args
    perm_out = np.empty(dh.ndofs, dtype=np.int64)
    if isinstance(order, DofOrder.FieldWise):
        tb = None if order.target_blocks is None else _i64(order.target_blocks)
        L.call("fb2_dh_renumber", dh.h, order.kind, _ptr(tb, C.c_int64) if tb is not None else None,
               0 if tb is None else len(tb), None, _ptr(perm_out, C.c_int64))
    else:
        perm = _i64(order)
        assert perm.shape == (dh.ndofs,), "input vector is not a permutation of length ndofs(dh)"
        L.call("fb2_dh_renumber", dh.h, L.ORDER_PERMUTATION, None, 0, _ptr(perm, C.c_int64), _ptr(perm_out, C.c_int64))
    if ch is not None:
        L.call("fb2_ch_renumber", ch.h, _ptr(perm_out, C.c_int64))
    return perm_out


def add_(obj, *args):
    """add!(dh, name, ip)  or  add!(ch, Dirichlet / AffineConstraint / PeriodicDirichlet)"""
    if isinstance(obj, DofHandler):
        name, ip = args
        assert obj.h is None, "DofHandler already closed"
        assert ip.refshape == obj.grid.celltype
        obj.field_names.append(name)
        obj.field_ips.append(ip)
        return obj
    if isinstance(obj, ConstraintHandler):
        if isinstance(args[0], AffineConstraint):
            return obj._add_affine(args[0])
        if isinstance(args[0], PeriodicDirichlet):
            return obj._add_periodic(args[0])
        return obj._add(args[0])
    raise TypeError(type(obj))


def close_(obj):
    """close!(dh) / close!(ch)"""
    if isinstance(obj, DofHandler):
        obj.h = C.c_void_p()
        L.call("fb2_dh_close", obj.grid.h, len(obj.field_ips), obj._fields(), C.byref(obj.h))
        obj._info()
        return obj
    if isinstance(obj, ConstraintHandler):
        return obj._close()
    raise TypeError(type(obj))


def ndofs(dh):
    return dh.ndofs


def ndofs_per_cell(dh):
    return dh.ndofs_per_cell


def celldofs(dh, i):
    """celldofs(dh, i), i 1-based"""
    return dh.cell_dofs[i - 1]


def dof_range(dh, name):
    a, b = C.c_int(), C.c_int()
    L.call("fb2_dh_dof_range", dh.h, dh.field_names.index(name), C.byref(a), C.byref(b))
    return range(a.value, b.value + 1)


# ---- matrix -----------------------------------------------------------------------------------------------
class B200Matrix:
    """Device-resident SparseMatrixCSC{Float64,Int}: pattern handle + nzval CUDA tensor."""

    def __init__(self, dh, handle):
        self.dh, self.h = dh, handle
        n, nnz = C.c_int64(), C.c_int64()
        L.call("fb2_pattern_info", self.h, C.byref(n), C.byref(nnz))
        self.n, self.nnz = n.value, nnz.value
        self.nzval = dh.grid.ctx.zeros(self.nnz)
        self._colptr = self._rowval = None
        self._asm = {}       # native assemblers of this matrix, one per CellValues: id(cv) -> (handle, cv)

    def _export(self):
        if self._colptr is None:
            self._colptr = np.empty(self.n + 1, dtype=np.int64)
            self._rowval = np.empty(self.nnz, dtype=np.int64)
            L.call("fb2_pattern_export", self.h, _ptr(self._colptr, C.c_int64), _ptr(self._rowval, C.c_int64))

    @property
    def colptr(self):
        self._export()
        return self._colptr

    @property
    def rowval(self):
        self._export()
        return self._rowval

    def tocsc(self):
        import scipy.sparse as sp
        self.dh.grid.ctx.synchronize()
        return sp.csc_matrix((self.nzval.cpu().numpy(), self.rowval - 1, self.colptr - 1), shape=(self.n, self.n))

    def __del__(self):
        if _destroy is not None:      # module globals are already gone at interpreter shutdown
            try:
                alive = getattr(self, "h", None) and all(getattr(p, "h", None) for p in _chain(self, "dh"))
                for h, _ in getattr(self, "_asm", {}).values():
                    if alive:
                        L.lib.fb2_assembler_destroy(h)
                self._asm = {}
            except Exception:
                pass
            _destroy(self, "fb2_pattern_destroy", _chain(self, "dh"))


def allocate_matrix(dh, colptr=None, rowval=None, ch=None):
    """allocate_matrix(dh) (pattern built on the device); allocate_matrix(dh, ch=ch): the condensed pattern for a closed
    ConstraintHandler with affine / periodic constraints; with colptr/rowval: adopt the reference's pattern."""
    h = C.c_void_p()
    if isinstance(colptr, ConstraintHandler):
        ch, colptr = colptr, None
    if ch is not None:
        L.call("fb2_pattern_create_condensed", dh.h, ch.h, C.byref(h))
    elif colptr is None:
        L.call("fb2_pattern_create", dh.h, C.byref(h))
    else:
        cp, rv = _i64(colptr), _i64(rowval)
        L.call("fb2_pattern_from_host", dh.h, _ptr(cp, C.c_int64), _ptr(rv, C.c_int64), C.byref(h))
    return B200Matrix(dh, h)


# ---- CellValues ----------------------------------------------------------------------------------------------
class CellValues:
    """CellValues(qr, ip[, ip_geo])"""

    def __init__(self, qr, ip, ip_geo=None, ctx=None, tables=None):
        self.ctx = ctx or default_context()
        self.qr, self.ip = qr, ip
        self.h = C.c_void_p()
        if tables is None:
            geo_order = ip_geo.order if ip_geo is not None else 1
            L.call("fb2_cellvalues_create", self.ctx.h, ip.refshape, qr.order, ip.order, ip.vdim, geo_order, C.byref(self.h))
        else:  # arrays-in: N (nq,n), dNdxi (nq,n,rdim), M (nq,ngeo), dMdxi (nq,ngeo,rdim), w (nq)
            N, dN, M, dM, w = (_f64(t) for t in tables)
            L.call("fb2_cellvalues_from_tables", self.ctx.h, ip.refshape, N.shape[0], N.shape[1], ip.vdim, M.shape[1],
                   _ptr(N, C.c_double), _ptr(dN, C.c_double), _ptr(M, C.c_double), _ptr(dM, C.c_double),
                   _ptr(w, C.c_double), C.byref(self.h))
        nq, nb, vd, ng, rd = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.call("fb2_cellvalues_info", self.h, C.byref(nq), C.byref(nb), C.byref(vd), C.byref(ng), C.byref(rd))
        self.nq, self.nbase_scalar, self.vdim, self.ngeo, self.rdim = nq.value, nb.value, vd.value, ng.value, rd.value

    def tables(self):
        N = np.empty((self.nq, self.nbase_scalar))
        dN = np.empty((self.nq, self.nbase_scalar, self.rdim))
        M = np.empty((self.nq, self.ngeo))
        dM = np.empty((self.nq, self.ngeo, self.rdim))
        w = np.empty(self.nq)
        pts = np.empty((self.nq, self.rdim))
        L.call("fb2_cellvalues_export", self.h, _ptr(N, C.c_double), _ptr(dN, C.c_double), _ptr(M, C.c_double),
               _ptr(dM, C.c_double), _ptr(w, C.c_double), _ptr(pts, C.c_double))
        return dict(N=N, dNdxi=dN, M=M, dMdxi=dM, w=w, points=pts)

    def __del__(self):
        if _destroy is not None:      # module globals are already gone at interpreter shutdown
            _destroy(self, "fb2_cellvalues_destroy", (getattr(self, "ctx", None),))


def reinit_(cv, grid, cells=None):
    """reinit!(cv, cell) for a batch of cells (1-based ids; None = all): returns CUDA tensors
    dNdx (n, nq, nbase, dim) and detJdV (n, nq) -- shape_gradient(cv, q, i) and getdetJdV(cv, q) of every cell."""
    torch = _torch()
    n = grid.ncells if cells is None else len(cells)
    dev = f"cuda:{grid.ctx.device}"
    dNdx = torch.empty((n, cv.nq, cv.nbase_scalar, cv.rdim), dtype=torch.float64, device=dev)
    dO = torch.empty((n, cv.nq), dtype=torch.float64, device=dev)
    ids = None if cells is None else _i64(cells)
    L.call("fb2_reinit_cells", cv.h, grid.h, _ptr(ids, C.c_int64) if ids is not None else None, n,
           C.c_void_p(dNdx.data_ptr()), C.c_void_p(dO.data_ptr()))
    grid.ctx.synchronize()
    return dNdx, dO


def spatial_coordinates_(cv, grid, cells=None):
    """spatial_coordinate(cv, q, x) of every quadrature point of a batch of cells: CUDA tensor (n, nq, sdim)."""
    torch = _torch()
    n = grid.ncells if cells is None else len(cells)
    x = torch.empty((n, cv.nq, grid.sdim), dtype=torch.float64, device=f"cuda:{grid.ctx.device}")
    ids = None if cells is None else _i64(cells)
    L.call("fb2_spatial_coordinates", cv.h, grid.h, _ptr(ids, C.c_int64) if ids is not None else None, n, C.c_void_p(x.data_ptr()))
    grid.ctx.synchronize()
    return x


class ReinitCellValues:
    """`cv` after reinit!(cv, cell) for a whole batch of cells: the accessors of src/FEValues/common_values.jl and CellValues.jl
    (shape_value :138ff, shape_gradient, shape_symmetric_gradient, shape_divergence, getdetJdV, spatial_coordinate :363-372,
    function_value / function_gradient :177-227) with 1-based q and i as in the reference.  Values that depend on the cell come
    back as CUDA tensors whose first axis is the cell of the batch; they are views of the buffers fb2_reinit_cells and
    fb2_spatial_coordinates filled on the device."""

    def __init__(self, cv, grid, cells=None):
        self.cv, self.grid, self.cells = cv, grid, cells
        self.dNdx, self.detJdV = reinit_(cv, grid, cells)
        self.N = _torch().from_numpy(cv.tables()["N"]).to(self.dNdx.device)      # (nq, nbase)
        self._x = None

    def getnquadpoints(self):
        return self.cv.nq

    def getnbasefunctions(self):
        return self.cv.nbase_scalar * self.cv.vdim

    def _split(self, i):
        assert 1 <= i <= self.getnbasefunctions(), "base function index out of range"
        return (i - 1) // self.cv.vdim, (i - 1) % self.cv.vdim

    def shape_value(self, q, i):
        """N_i(xi_q): a float for scalar interpolations, a (vdim,) tensor N_a e_c for vectorised ones (i = (a-1) vdim + c)"""
        a, c = self._split(i)
        v = self.N[q - 1, a]
        if self.cv.vdim == 1:
            return float(v)
        out = _torch().zeros(self.cv.vdim, dtype=_torch().float64, device=self.N.device)
        out[c] = v
        return out

    def shape_gradient(self, q, i):
        """(n, dim) for scalar interpolations; (n, vdim, dim) with row c = grad N_a for vectorised ones"""
        a, c = self._split(i)
        g = self.dNdx[:, q - 1, a, :]
        if self.cv.vdim == 1:
            return g
        out = _torch().zeros((g.shape[0], self.cv.vdim, g.shape[1]), dtype=g.dtype, device=g.device)
        out[:, c, :] = g
        return out

    def shape_symmetric_gradient(self, q, i):
        g = self.shape_gradient(q, i)
        return 0.5 * (g + g.transpose(1, 2))

    def shape_divergence(self, q, i):
        a, c = self._split(i)
        return self.dNdx[:, q - 1, a, c] if self.cv.vdim > 1 else self.dNdx[:, q - 1, a, :].sum(dim=1)

    def getdetJdV(self, q):
        return self.detJdV[:, q - 1]

    def spatial_coordinate(self, q):
        if self._x is None:
            self._x = spatial_coordinates_(self.cv, self.grid, self.cells)
        return self._x[:, q - 1, :]

    def function_value(self, q, ue):
        """ue: (n, ndofs_per_cell) cell-local dof values -> (n,) or (n, vdim)"""
        u = ue.reshape(ue.shape[0], self.cv.nbase_scalar, self.cv.vdim)
        v = _torch().einsum("a,nac->nc", self.N[q - 1], u)
        return v[:, 0] if self.cv.vdim == 1 else v

    def function_gradient(self, q, ue):
        u = ue.reshape(ue.shape[0], self.cv.nbase_scalar, self.cv.vdim)
        g = _torch().einsum("nac,nad->ncd", u, self.dNdx[:, q - 1])
        return g[:, 0, :] if self.cv.vdim == 1 else g


    def function_symmetric_gradient(self, q, ue):
        g = self.function_gradient(q, ue)
        return 0.5 * (g + g.transpose(1, 2))

    def function_divergence(self, q, ue):
        g = self.function_gradient(q, ue)
        return g.sum(dim=1) if self.cv.vdim == 1 else g.diagonal(dim1=1, dim2=2).sum(dim=1)


def reinit_batch(cv, grid, cells=None):
    return ReinitCellValues(cv, grid, cells)


def function_values_(cv, dh, u, gradients=True):
    """function_value(cv, q, ue) and function_gradient(cv, q, ue) at every quadrature point of every cell:
    CUDA tensors (ncells, nq, vdim) and (ncells, nq, vdim, dim)."""
    torch = _torch()
    g = dh.grid
    dev = f"cuda:{g.ctx.device}"
    vals = torch.empty((g.ncells, cv.nq, cv.vdim), dtype=torch.float64, device=dev)
    grads = torch.empty((g.ncells, cv.nq, cv.vdim, cv.rdim), dtype=torch.float64, device=dev) if gradients else None
    L.call("fb2_function_values", cv.h, dh.h, C.c_void_p(u.data_ptr()), C.c_void_p(vals.data_ptr()),
           C.c_void_p(grads.data_ptr()) if grads is not None else None)
    g.ctx.synchronize()
    return (vals, grads) if gradients else vals


# ---- FacetValues and the Neumann / traction facet loop ---------------------------------------------------
class FacetQuadratureRule:
    """FacetQuadratureRule{refshape}(order) (src/Quadrature/quadrature.jl:205-238)"""

    def __init__(self, refshape, order):
        self.refshape, self.order = refshape, int(order)


class FacetValues:
    """FacetValues(fqr, ip[, ip_geo]) (src/FEValues/FacetValues.jl:39-88)"""

    def __init__(self, fqr, ip, ip_geo=None, ctx=None):
        self.ctx = ctx or default_context()
        self.fqr, self.ip = fqr, ip
        self.h = C.c_void_p()
        geo_order = ip_geo.order if ip_geo is not None else 1
        L.call("fb2_facetvalues_create", self.ctx.h, ip.refshape, fqr.order, ip.order, ip.vdim, geo_order, C.byref(self.h))
        nf, nq, nb, vd, rd = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.call("fb2_facetvalues_info", self.h, C.byref(nf), C.byref(nq), C.byref(nb), C.byref(vd), C.byref(rd))
        self.nfacets, self.nq, self.nbase_scalar, self.vdim, self.rdim = nf.value, nq.value, nb.value, vd.value, rd.value

    def tables(self):
        w = np.empty((self.nfacets, self.nq))
        pts = np.empty((self.nfacets, self.nq, self.rdim))
        N = np.empty((self.nfacets, self.nq, self.nbase_scalar))
        L.call("fb2_facetvalues_export", self.h, _ptr(w, C.c_double), _ptr(pts, C.c_double), _ptr(N, C.c_double))
        return dict(w=w, points=pts, N=N)

    def __del__(self):
        if _destroy is not None:      # module globals are already gone at interpreter shutdown
            _destroy(self, "fb2_facetvalues_destroy", (getattr(self, "ctx", None),))


class FacetSet:
    """A set of FacetIndex (cell, local facet) pairs, 1-based: getfacetset(grid, name) or a union of several."""

    def __init__(self, grid, pairs):
        self.grid = grid
        self.pairs = _i64(np.asarray(pairs).reshape(-1, 2))
        self.h = C.c_void_p()
        L.call("fb2_facetset_create", grid.h, _ptr(self.pairs, C.c_int64), self.pairs.shape[0], C.byref(self.h))

    def __del__(self):
        if _destroy is not None:      # module globals are already gone at interpreter shutdown
            _destroy(self, "fb2_facetset_destroy", (getattr(self, "grid", None),))


def assemble_facets_(f, dh, fv, facetset, kind, params):
    """The facet loop `for (cell, facet) in set: reinit!(fv, cell, facet); ...; assemble!(f, celldofs, fe)`
    (hyperelasticity.jl:278-291).  kind: 'flux' (params q), 'traction' (params t[vdim]) or 'normal_traction' (params p:
    fe += p n N dGamma; the tutorial's `ge[i] -= (dui . tn n) dGamma` is p = -tn).  f: device vector (torch / pointer)."""
    if not isinstance(facetset, FacetSet):
        facetset = FacetSet(dh.grid, facetset)
    k = {"flux": L.FACET_FLUX, "traction": L.FACET_TRACTION, "normal_traction": L.FACET_NORMAL_TRACTION}[kind]
    p = _f64(np.atleast_1d(np.asarray(params, dtype=np.float64)))
    L.call("fb2_assemble_facets", dh.h, fv.h, facetset.h, k, _ptr(p, C.c_double), int(p.size), C.c_void_p(f.data_ptr()))
    return f


# ---- element menu -----------------------------------------------------------------------------------------------
class HeatElement:
    """Ke = int k grad(Ni).grad(Nj), fe = int source Ni  (heat_equation.jl:143-164)"""
    elem_id = L.ELEM_HEAT

    def __init__(self, k=1.0, source=1.0):
        self.params = L.HeatParams(float(k), float(source))


class MassElement:
    elem_id = L.ELEM_MASS

    def __init__(self, rho=1.0):
        self.params = L.MassParams(float(rho))


class ElasticityElement:
    """Ke = int eps_i : C : eps_j with isotropic C(lambda, mu), fe = int Ni . b  (threaded_assembly.jl:105-119)"""
    elem_id = L.ELEM_ELASTICITY

    def __init__(self, lam=None, mu=None, b=(0.0, 0.0, 0.0), E=None, nu=None):
        if lam is None:
            lam = E * nu / ((1 + nu) * (1 - 2 * nu))
            mu = E / (2 * (1 + nu))
        b = tuple(b) + (0.0,) * (3 - len(b))
        self.params = L.ElasticityParams(float(lam), float(mu), (C.c_double * 3)(*b))


class GeneralElasticityElement:
    """Ke = int grad(dN_i) : C : grad(N_j) for any stiffness tensor C with minor symmetries (orthotropic, anisotropic, ...),
    fe = int Ni . b -- the element routine as the reference writes it (linear_elasticity.jl:266-281, benchmark/helper.jl:249-262).
    C: (dim, dim, dim, dim) array = SymmetricTensor{4, dim}.  ElasticityElement is the isotropic fast path."""
    elem_id = L.ELEM_ELASTICITY_GENERAL

    def __init__(self, C4, b=(0.0, 0.0, 0.0)):
        C4 = np.asarray(C4, dtype=np.float64)
        d = C4.shape[0]
        assert C4.shape == (d,) * 4 and d in (2, 3)
        full = np.zeros((3, 3, 3, 3))
        full[:d, :d, :d, :d] = C4
        b = tuple(b) + (0.0,) * (3 - len(b))
        self.params = L.ElasticityGeneralParams((C.c_double * 81)(*full.ravel()), (C.c_double * 3)(*b))


class NeoHookeElement(ElasticityElement):
    """tangent + residual of the compressible Neo-Hooke model (hyperelasticity.jl:162-176,241-276)"""
    elem_id = L.ELEM_NEOHOOKE


class MultiFieldCellValues:
    """MultiFieldCellValues(qr, (u = ip_u, p = ip_p)) (src/FEValues/CellValues.jl:229-298): several interpolations evaluated on one
    quadrature rule and one geometric mapping; the fields are attributes like `cellvalues.u` / `cellvalues.p`."""

    def __init__(self, qr, ip_geo=None, ctx=None, **fields):
        self.qr, self.names = qr, list(fields)
        for name, ip in fields.items():
            setattr(self, name, CellValues(qr, ip, ip_geo, ctx=ctx))

    def __getitem__(self, name):
        return getattr(self, name)


# ---- assembler ------------------------------------------------------------------------------------------------------
class Assembler:
    """The assembler start_assemble returns.  The native assembler (the cell-local -> nzval map, src/assembler.jl:347-457) is
    cached on the matrix per CellValues, so calling start_assemble before every assembly -- as the reference does -- is free.
    `fillzero` is consumed by the first assemble_ / scatter_ call: start_assemble zeroes K and f ONCE (src/assembler.jl:
    287-291), every later call on the same assembler adds onto the result."""

    def __init__(self, K, f, fillzero=True, scatter="atomic"):
        self.K, self.f = K, f
        self.fillzero = fillzero
        self.scatter = scatter
        self.variant = 0

    def _handle(self, cv):
        cache = self.K._asm
        key = id(cv)
        if key not in cache:
            h = C.c_void_p()
            L.call("fb2_assembler_create", self.K.dh.h, self.K.h, cv.h if cv is not None else None, C.byref(h))
            cache[key] = (h, cv)
        return cache[key][0]

    def _opts(self):
        """options of the next native call; the pending zero fill is handed to that call and cleared"""
        o = L.AsmOpts(1 if self.fillzero else 0, L.SCATTER_COLORED if self.scatter == "colored" else L.SCATTER_ATOMIC,
                      self.variant, 0)
        self.fillzero = False
        return o

    def coloring(self, cv=None):
        h = self._handle(cv)
        nc = C.c_int()
        col = np.empty(self.K.dh.grid.ncells, dtype=np.int32)
        L.call("fb2_assembler_coloring", h, C.byref(nc), _ptr(col, C.c_int32))
        return nc.value, col


def start_assemble(K, f=None, fillzero=True, scatter="atomic"):
    """start_assemble(K, f; fillzero): K.nzval and f are zeroed once -- the fill is fused into the FIRST assemble_ call on the
    returned assembler (nothing may read K in between); later assemble_ calls on the same assembler accumulate, like
    repeated assemble! calls of the reference."""
    return Assembler(K, f, fillzero, scatter)


def assemble_(assembler, element, cv, u=None):
    """The whole `for cell in CellIterator(dh): reinit!, element routine, assemble!` loop on the GPU."""
    h = assembler._handle(cv)
    opts = assembler._opts()
    f = assembler.f
    L.call("fb2_assemble", h, element.elem_id, C.byref(element.params), C.sizeof(element.params),
           C.c_void_p(u.data_ptr()) if u is not None else None,
           C.c_void_p(assembler.K.nzval.data_ptr()), C.c_void_p(f.data_ptr()) if f is not None else None, C.byref(opts))
    return assembler


def assemble_mixed_up_(assembler, cellvalues, G, K_bulk, u="u", p="p"):
    """The cell loop of the incompressible-elasticity tutorial (assemble_up!, incompressible_elasticity.jl:266-293) for a
    two-field DofHandler: cellvalues = MultiFieldCellValues(qr, u=ip_u, p=ip_p); K_bulk may be inf.  The traction term is added
    with assemble_facets_ afterwards."""
    dh = assembler.K.dh
    h = assembler._handle(None)
    opts = assembler._opts()
    f = assembler.f
    invK = 0.0 if np.isinf(K_bulk) else 1.0 / float(K_bulk)
    L.call("fb2_assemble_mixed_up", h, cellvalues[u].h, cellvalues[p].h, dh.field_names.index(u), dh.field_names.index(p), float(G),
           invK, C.c_void_p(assembler.K.nzval.data_ptr()), C.c_void_p(f.data_ptr()) if f is not None else None, C.byref(opts))
    return assembler


def assemble_host(assembler, element, cv, nzval, f=None, u=None):
    """Same loop through host buffers (numpy float64): nzval/f are filled like SparseMatrixCSC.nzval / f."""
    h = assembler._handle(cv)
    opts = assembler._opts()
    L.call("fb2_assemble_host", h, element.elem_id, C.byref(element.params), C.sizeof(element.params),
           _ptr(u, C.c_double) if u is not None else None, _ptr(nzval, C.c_double),
           _ptr(f, C.c_double) if f is not None else None, C.byref(opts))


def assemble_host_streamed(assembler, element, cv, nzval, f=None, u=None, xyz=None):
    """assemble_host with the copies pipelined against the assembly (slabs of cells): xyz (nnodes, sdim) host coordinates
    of this step (optional), nzval / f host outputs.  Pinned buffers let the copies overlap."""
    h = assembler._handle(cv)
    opts = assembler._opts()
    L.call("fb2_assemble_host_streamed", h, element.elem_id, C.byref(element.params), C.sizeof(element.params),
           _ptr(xyz, C.c_double) if xyz is not None else None, _ptr(u, C.c_double) if u is not None else None,
           _ptr(nzval, C.c_double), _ptr(f, C.c_double) if f is not None else None, C.byref(opts))


def scatter_(assembler, Ke, fe=None):
    """assemble!(assembler, dofs, Ke, fe) for all cells at once from precomputed element matrices:
    Ke (ncells, n, n) with Ke[c, i, j], fe (ncells, n)."""
    h = assembler._handle(None)
    opts = assembler._opts()
    Kc = _f64(np.transpose(np.asarray(Ke), (0, 2, 1)))   # -> per cell column-major
    fp = _ptr(_f64(fe), C.c_double) if fe is not None else None
    f = assembler.f
    L.call("fb2_scatter_host", h, _ptr(Kc, C.c_double), fp, C.c_void_p(assembler.K.nzval.data_ptr()),
           C.c_void_p(f.data_ptr()) if f is not None else None, C.byref(opts))


def scatter_device_(assembler, Kes, fes=None):
    """assemble!(assembler, celldofs(cell), Ke, fe) for all cells from device-resident element matrices in the layout of
    ElementAssembly.assemble (CUDA tensors: Kes[c, j, i] = Ke_c[i, j], fes[c, i])."""
    h = assembler._handle(None)
    opts = assembler._opts()
    f = assembler.f
    L.call("fb2_scatter_device", h, C.c_void_p(Kes.data_ptr()), C.c_void_p(fes.data_ptr()) if fes is not None else None,
           C.c_void_p(assembler.K.nzval.data_ptr()), C.c_void_p(f.data_ptr()) if f is not None else None, C.byref(opts))
    return assembler


class ElementAssembly:
    """"Element assembly" (docs/src/literate-howto/gpu_assembly.jl:265-304): all element matrices are kept on the device,
    Kes (ncells, n, n) with Kes[c, j, i] = Ke_c[i, j] (column-major per cell) and fes (ncells, n); `mul` is the
    matrix-free operator y = sum_e P_e' Ke P_e x, `apply_local_` is apply_local! on every cell."""

    def __init__(self, dh, cv):
        self.dh, self.cv = dh, cv
        self.h = C.c_void_p()
        L.call("fb2_ea_create", dh.h, cv.h, C.byref(self.h))
        nc, n = C.c_int64(), C.c_int()
        L.call("fb2_ea_info", self.h, C.byref(nc), C.byref(n))
        self.ncells, self.n = nc.value, n.value

    def assemble(self, element, u=None, Kes=None, fes=None):
        ctx = self.dh.grid.ctx
        if Kes is None:
            Kes = ctx.zeros(self.ncells * self.n * self.n).view(self.ncells, self.n, self.n)
        if fes is None:
            fes = ctx.zeros(self.ncells * self.n).view(self.ncells, self.n)
        L.call("fb2_ea_assemble", self.h, element.elem_id, C.byref(element.params), C.sizeof(element.params),
               C.c_void_p(u.data_ptr()) if u is not None else None, C.c_void_p(Kes.data_ptr()), C.c_void_p(fes.data_ptr()))
        return Kes, fes

    def mul(self, Kes, x, out=None):
        y = out if out is not None else self.dh.grid.ctx.zeros(self.dh.ndofs)
        L.call("fb2_ea_mul", self.h, C.c_void_p(Kes.data_ptr()), C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()))
        return y

    def diag(self, Kes, out=None):
        d = out if out is not None else self.dh.grid.ctx.zeros(self.dh.ndofs)
        L.call("fb2_ea_diag", self.h, C.c_void_p(Kes.data_ptr()), C.c_void_p(d.data_ptr()))
        return d

    def rhs(self, fes, out=None):
        """f = sum_e P_e' fe"""
        f = out if out is not None else self.dh.grid.ctx.zeros(self.dh.ndofs)
        L.call("fb2_ea_rhs", self.h, C.c_void_p(fes.data_ptr()), C.c_void_p(f.data_ptr()))
        return f

    def cg_(self, x, Kes, b, reltol=None, abstol=0.0, maxiter=None, jacobi=False):
        """IterativeSolvers.cg!(x, A, b) with the matrix-free operator A = (this, Kes); returns (iterations, residual norm)."""
        reltol = float(np.sqrt(np.finfo(np.float64).eps)) if reltol is None else float(reltol)
        maxiter = self.dh.ndofs if maxiter is None else int(maxiter)
        it, res = C.c_int(), C.c_double()
        L.call("fb2_ea_cg", self.h, C.c_void_p(Kes.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), reltol,
               float(abstol), maxiter, 1 if jacobi else 0, C.byref(it), C.byref(res))
        return it.value, res.value

    def apply_local_(self, Kes, fes, ch, applyzero=False):
        L.call("fb2_ea_apply_local", self.h, ch.h, C.c_void_p(Kes.data_ptr()),
               C.c_void_p(fes.data_ptr()) if fes is not None else None, 1 if applyzero else 0)

    def __del__(self):
        if _destroy is not None:
            _destroy(self, "fb2_ea_destroy", [self.cv] + _chain(self, "dh"))


def apply_assemble_(assembler, ch, element, cv, u=None, applyzero=False, ea=None):
    """The cell loop with apply_assemble!(assembler, ch, celldofs(cell), Ke, fe; apply_zero) in place of assemble!
    (src/assembler.jl:491-503): the Dirichlet conditions are applied element by element, no global apply! afterwards.
    `ea`: an ElementAssembly to reuse between calls (its scratch holds all element matrices)."""
    ea = ea if ea is not None else ElementAssembly(assembler.K.dh, cv)
    h = assembler._handle(cv)
    opts = assembler._opts()
    L.call("fb2_apply_assemble", h, ea.h, ch.h, element.elem_id, C.byref(element.params), C.sizeof(element.params),
           C.c_void_p(u.data_ptr()) if u is not None else None, C.c_void_p(assembler.K.nzval.data_ptr()),
           C.c_void_p(assembler.f.data_ptr()), 1 if applyzero else 0, C.byref(opts))
    return ea


def last_kernel():
    """name of the kernel family the last assemble call launched for the cell loop (k_march_hex, k_cell_scalar, ...)"""
    return L.lib.fb2_last_kernel().decode()


def finish_assemble(assembler):
    assembler.K.dh.grid.ctx.synchronize()
    return assembler.K, assembler.f


# ---- constraints ---------------------------------------------------------------------------------------------------------
class Dirichlet:
    """Dirichlet(field, set, f, components): set = getfacetset(...) pairs (kind 'facet'/'face'/'edge'/'vertex')
    or node ids (kind 'node'); f(x, t) -> value(s)."""

    def __init__(self, field, entities, f, components=None, kind="facet"):
        self.field, self.f, self.kind = field, f, kind
        self.entities = _i64(entities)
        self.components = None if components is None else [int(c) for c in np.atleast_1d(components)]


_KIND = {"facet": L.BC_FACET, "face": L.BC_FACE, "edge": L.BC_EDGE, "vertex": L.BC_VERTEX, "node": L.BC_NODE}


class ConstraintHandler:
    def __init__(self, dh):
        self.dh = dh
        self.h = C.c_void_p()
        L.call("fb2_ch_create", dh.h, C.byref(self.h))
        self.dbcs = []
        self.closed = False

    @classmethod
    def from_arrays(cls, dh, prescribed_dofs, inhomogeneities):
        ch = cls.__new__(cls)
        ch.dh, ch.dbcs, ch.closed = dh, [], True
        ch.h = C.c_void_p()
        p, v = _i64(prescribed_dofs), _f64(inhomogeneities)
        L.call("fb2_ch_from_host", dh.h, len(p), _ptr(p, C.c_int64), _ptr(v, C.c_double), C.byref(ch.h))
        return ch

    def set_inhomogeneities(self, values):
        """update!(ch, t) in arrays-in mode: the reference's ch.inhomogeneities (order of prescribed_dofs)"""
        v = _f64(values)
        L.call("fb2_ch_set_inhomogeneities", self.h, len(v), _ptr(v, C.c_double))

    def _add(self, dbc):
        comps = dbc.components or []
        carr = (C.c_int * max(len(comps), 1))(*comps)
        ibc = C.c_int()
        n = dbc.entities.shape[0]
        L.call("fb2_ch_add_dirichlet", self.h, self.dh.field_names.index(dbc.field), _KIND[dbc.kind], n,
               _ptr(dbc.entities, C.c_int64), len(comps), carr, C.byref(ibc))
        ncomp = len(comps) if comps else self.dh.field_ips[self.dh.field_names.index(dbc.field)].vdim
        self.dbcs.append((ibc.value, dbc, ncomp))
        return self

    def _add_affine(self, ac):
        m = _i64([d for d, _ in ac.entries])
        v = _f64([c for _, c in ac.entries])
        L.call("fb2_ch_add_affine", self.h, int(ac.constrained_dof), len(m), _ptr(m, C.c_int64) if len(m) else None,
               _ptr(v, C.c_double) if len(m) else None, float(ac.b))
        return self

    def _add_periodic(self, pd):
        mset, iset = _i64(pd.face_map[0]), _i64(pd.face_map[1])
        comps = pd.components or []
        carr = (C.c_int * max(len(comps), 1))(*comps)
        L.call("fb2_ch_add_periodic", self.h, self.dh.field_names.index(pd.field), mset.shape[0], _ptr(mset, C.c_int64),
               iset.shape[0], _ptr(iset, C.c_int64), len(comps), carr)
        return self

    @property
    def dofcoefficients(self):
        """ch.dofcoefficients: per prescribed dof None or [(master, coeff), ...] (1-based masters)"""
        n, nt = C.c_int64(), C.c_int64()
        L.call("fb2_ch_info", self.h, C.byref(n))
        L.call("fb2_ch_affine_export", self.h, C.byref(nt), None, None, None)
        ptr, m, v = np.empty(n.value + 1, dtype=np.int64), np.empty(nt.value, dtype=np.int64), np.empty(nt.value)
        L.call("fb2_ch_affine_export", self.h, C.byref(nt), _ptr(ptr, C.c_int64), _ptr(m, C.c_int64) if nt.value else None,
               _ptr(v, C.c_double) if nt.value else None)
        return [list(zip(m[ptr[i]:ptr[i + 1]].tolist(), v[ptr[i]:ptr[i + 1]].tolist())) if ptr[i + 1] > ptr[i] else None
                for i in range(n.value)]

    def _close(self):
        L.call("fb2_ch_close", self.h)
        self.closed = True
        update_(self, 0.0)
        return self

    @property
    def prescribed_dofs(self):
        n = C.c_int64()
        L.call("fb2_ch_info", self.h, C.byref(n))
        out = np.empty(n.value, dtype=np.int64)
        L.call("fb2_ch_export", self.h, _ptr(out, C.c_int64), None)
        return out

    @property
    def inhomogeneities(self):
        n = C.c_int64()
        L.call("fb2_ch_info", self.h, C.byref(n))
        out = np.empty(n.value, dtype=np.float64)
        L.call("fb2_ch_export", self.h, None, _ptr(out, C.c_double))
        return out

    def __del__(self):
        if _destroy is not None:      # module globals are already gone at interpreter shutdown
            _destroy(self, "fb2_ch_destroy", _chain(self, "dh"))


class AffineConstraint:
    """AffineConstraint(constrained_dof, [master => coeff, ...], b): u_dof = sum coeff u_master + b
    (src/Dofs/ConstraintHandler.jl:114-131); entries = [(master_dof, coeff), ...], dofs 1-based"""

    def __init__(self, constrained_dof, entries, b=0.0):
        self.constrained_dof, self.entries, self.b = int(constrained_dof), [(int(d), float(c)) for d, c in entries], float(b)


def collect_periodic_facets(grid, mset, iset):
    """collect_periodic_facets(grid, mirror_set, image_set) for facet sets that are translates of each other: the pair of
    facet sets; the dof pairing (by position) happens in add!(ch, PeriodicDirichlet(...))."""
    m = getfacetset(grid, mset) if isinstance(mset, str) else mset
    i = getfacetset(grid, iset) if isinstance(iset, str) else iset
    return (_i64(m), _i64(i))


class PeriodicDirichlet:
    """PeriodicDirichlet(field, face_map[, components]): the dofs on the mirror facets are constrained to those on the image
    facets, u_mirror = u_image (src/Dofs/ConstraintHandler.jl:1032-1300)"""

    def __init__(self, field, face_map, components=None):
        self.field, self.face_map, self.components = field, face_map, list(components) if components else None


def update_(ch, t=0.0):
    """update!(ch, t): evaluate every Dirichlet function at its dof locations (in the reference's order)."""
    sdim = ch.dh.grid.sdim
    for ibc, dbc, ncomp in ch.dbcs:
        n = C.c_int64()
        L.call("fb2_ch_bc_points", ch.h, ibc, C.byref(n), None)
        x = np.empty((n.value, sdim))
        L.call("fb2_ch_bc_points", ch.h, ibc, C.byref(n), _ptr(x, C.c_double))
        vals = np.empty((n.value, ncomp))
        for k in range(n.value):
            vals[k] = np.atleast_1d(dbc.f(x[k], t))
        L.call("fb2_ch_bc_set_values", ch.h, ibc, n.value, _ptr(vals, C.c_double))


def apply_(K, f=None, ch=None, applyzero=False):
    """apply!(K, f, ch) / apply!(u, ch) (first argument a CUDA tensor)."""
    if not isinstance(K, B200Matrix):   # apply!(u, ch)
        u, ch = K, f if ch is None else ch
        L.call("fb2_apply_vector", ch.h, C.c_void_p(u.data_ptr()), 1 if applyzero else 0)
        return None
    m = C.c_double()
    L.call("fb2_apply", ch.h, K.h, C.c_void_p(K.nzval.data_ptr()), C.c_void_p(f.data_ptr()) if f is not None else None,
           1 if applyzero else 0, C.byref(m))
    return m.value


class RHSData:
    """get_rhs_data(ch, A) (src/Dofs/ConstraintHandler.jl:191-208): mean diagonal and prescribed columns of A before apply!"""

    def __init__(self, ch, K):
        self.ch, self.K = ch, K
        self.h = C.c_void_p()
        L.call("fb2_rhsdata_create", ch.h, K.h, C.c_void_p(K.nzval.data_ptr()), C.byref(self.h))
        m, np_, ns = C.c_double(), C.c_int64(), C.c_int64()
        L.call("fb2_rhsdata_info", self.h, C.byref(m), C.byref(np_), C.byref(ns))
        self.m, self.nprescribed, self.nstored = m.value, np_.value, ns.value

    def __del__(self):
        if _destroy is not None:
            _destroy(self, "fb2_rhsdata_destroy", _chain(self.K, "dh"))


def get_rhs_data(ch, K):
    return RHSData(ch, K)


def apply_rhs_(data, f, ch, applyzero=False):
    """apply_rhs!(data, f, ch, applyzero) (src/Dofs/ConstraintHandler.jl:217-240): boundary conditions onto a new right-hand side"""
    L.call("fb2_apply_rhs", data.h, C.c_void_p(f.data_ptr()), ch.h, 1 if applyzero else 0)
    return f


def apply_zero_(K, f=None, ch=None):
    return apply_(K, f, ch, applyzero=True)


# ---- the step after the path: SpMV / CSR values / CG on the device -------------------------------------------------
def spmv(K, x, transpose=False, out=None):
    """y = K x (or K' x) with K's values where the assembly left them (CUDA tensors in, CUDA tensor out)."""
    y = out if out is not None else K.dh.grid.ctx.zeros(K.n)
    L.call("fb2_spmv", K.h, C.c_void_p(K.nzval.data_ptr()), C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), 1 if transpose else 0)
    return y


def csr_values(K):
    """nzval of K in CSR order; with rowptr = K.colptr and colval = K.rowval (structurally symmetric patterns)."""
    out = K.dh.grid.ctx.zeros(K.nnz)
    L.call("fb2_csr_values", K.h, C.c_void_p(K.nzval.data_ptr()), C.c_void_p(out.data_ptr()))
    return out


def cg_(x, K, b, reltol=None, abstol=0.0, maxiter=None, jacobi=False, symmetric=True):
    """IterativeSolvers.cg!(x, K, b; reltol = sqrt(eps), abstol = 0, maxiter = n) on the device.
    Returns (iterations, final residual norm); x is updated in place."""
    reltol = float(np.sqrt(np.finfo(np.float64).eps)) if reltol is None else float(reltol)
    maxiter = K.n if maxiter is None else int(maxiter)
    it, rn = C.c_int(), C.c_double()
    L.call("fb2_cg", K.h, C.c_void_p(K.nzval.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()),
           reltol, float(abstol), maxiter, 1 if jacobi else 0, 1 if symmetric else 0, C.byref(it), C.byref(rn))
    return it.value, rn.value


# ---- partitioned multi-GPU assembly (no counterpart in the reference) ---------------------------------------------
class Partition:
    """Rank `rank` of `nparts` of a global DofHandler: own + halo cells as a local problem, column ownership and
    per-peer interface exchange lists (see csrc/partition.cu).  `gdh` may live on a host-only context."""

    def __init__(self, gdh, nparts, rank, dims=None, cell_owner=None, metis=False):
        """dims: px, py, pz block layout for generate_grid input (None = automatic); cell_owner: any partitioner's
        cell -> rank array (0-based ranks) instead of the block layout; metis=True: METIS_PartMeshDual inside the library."""
        self.gdh, self.nparts, self.rank = gdh, int(nparts), int(rank)
        self.h = C.c_void_p()
        if metis:
            L.call("fb2_partition_create_metis", gdh.h, self.nparts, self.rank, C.byref(self.h))
        elif cell_owner is not None:
            own = np.ascontiguousarray(cell_owner, dtype=np.int32)
            assert own.shape == (gdh.grid.ncells,)
            L.call("fb2_partition_create_from_owners", gdh.h, self.nparts, self.rank, _ptr(own, C.c_int32), C.byref(self.h))
        else:
            darr = (C.c_int * 3)(*(list(dims) + [1, 1, 1])[:3]) if dims is not None else None
            L.call("fb2_partition_create", gdh.h, self.nparts, self.rank, darr, C.byref(self.h))
        self._field_names, self._field_ips = list(gdh.field_names), list(gdh.field_ips)
        self._export()

    @classmethod
    def generated(cls, nel, ip, nparts, rank, dims=None, left=None, right=None, perturb=0.0, name="u", host_ctx=None):
        """The plan Partition(close!(DofHandler(generate_grid(Hexahedron, nel, left, right).perturb(perturb)) + field `name` =
        Lagrange{RefHexahedron,1}^vdim), nparts, rank, dims) would give, built rank-locally from closed forms: no global grid,
        no global DofHandler (fb2_partition_create_generated)."""
        assert ip.refshape == Hexahedron and ip.order == 1
        self = cls.__new__(cls)
        self.gdh, self.nparts, self.rank = None, int(nparts), int(rank)
        self._hctx = host_ctx or Context(-1)
        self.h = C.c_void_p()
        nel = _i64(nel)
        lp = _ptr(_f64(left), C.c_double) if left is not None else None
        rp = _ptr(_f64(right), C.c_double) if right is not None else None
        darr = (C.c_int * 3)(*(list(dims) + [1, 1, 1])[:3]) if dims is not None else None
        L.call("fb2_partition_create_generated", self._hctx.h, _ptr(nel, C.c_int64), lp, rp, float(perturb), int(ip.vdim),
               self.nparts, self.rank, darr, C.byref(self.h))
        self.ncells_global = int(nel[0] * nel[1] * nel[2])
        self._field_names, self._field_ips = [name], [ip]
        self._export()
        return self

    def _export(self):
        v = [C.c_int64() for _ in range(5)]
        L.call("fb2_partition_info", self.h, *[C.byref(x) for x in v])
        self.ncells_local, self.ncells_own, self.nnodes_local, self.ndofs_local, self.ndofs_owned = (x.value for x in v)
        self.cells_global = np.empty(self.ncells_local, dtype=np.int64)
        self.cell_is_own = np.empty(self.ncells_local, dtype=np.uint8)
        self.l2g_node = np.empty(self.nnodes_local, dtype=np.int64)
        self.l2g_dof = np.empty(self.ndofs_local, dtype=np.int64)
        self.dof_owner = np.empty(self.ndofs_local, dtype=np.int32)
        L.call("fb2_partition_export", self.h, _ptr(self.cells_global, C.c_int64), _ptr(self.cell_is_own, C.c_uint8),
               _ptr(self.l2g_node, C.c_int64), _ptr(self.l2g_dof, C.c_int64), _ptr(self.dof_owner, C.c_int32))
        self._asm = None

    def local_problem(self, ctx):
        """(grid, dh) of the local sub-problem on `ctx` (local numbering)."""
        gh, dhh = C.c_void_p(), C.c_void_p()
        L.call("fb2_partition_local_grid", self.h, ctx.h, C.byref(gh))
        grid = Grid(ctx, gh)
        L.call("fb2_partition_local_dh", self.h, grid.h, C.byref(dhh))
        dh = DofHandler(grid)
        dh.field_names, dh.field_ips = list(self._field_names), list(self._field_ips)
        dh.h = dhh
        dh._info()
        return grid, dh

    def peer_counts(self, peer):
        v = [C.c_int64() for _ in range(4)]
        L.call("fb2_partition_peer_counts", self.h, int(peer), *[C.byref(x) for x in v])
        return tuple(x.value for x in v)      # nz_send, f_send, nz_recv, f_recv

    def peer_lists(self, peer):
        ns, fs, nr, fr = self.peer_counts(peer)
        a = [np.empty(n, dtype=np.int32) for n in (ns, ns, nr, nr, fs, fr)]
        L.call("fb2_partition_peer_lists", self.h, int(peer), *[_ptr(x, C.c_int32) for x in a])
        return dict(send_rows=a[0], send_cols=a[1], recv_rows=a[2], recv_cols=a[3], send_f=a[4], recv_f=a[5])

    def bind(self, assembler, cv):
        assembler._accumulate = not assembler.fillzero
        self._asm = assembler
        self._cv = cv
        L.call("fb2_partition_bind", self.h, assembler._handle(cv))

    def pack(self, peer, K, f, out):
        L.call("fb2_partition_pack", self.h, int(peer), C.c_void_p(K.nzval.data_ptr()),
               C.c_void_p(f.data_ptr()) if f is not None else None, C.c_void_p(out.data_ptr()))

    def unpack_add(self, peer, buf, K, f):
        L.call("fb2_partition_unpack_add", self.h, int(peer), C.c_void_p(buf.data_ptr()), C.c_void_p(K.nzval.data_ptr()),
               C.c_void_p(f.data_ptr()) if f is not None else None)

    def mask_unowned(self, K, f):
        L.call("fb2_partition_mask_unowned", self.h, C.c_void_p(K.nzval.data_ptr()),
               C.c_void_p(f.data_ptr()) if f is not None else None)

    def assemble_(self, element, mode="exchange", u=None):
        """assemble! for this rank: mode 'exchange' (own cells + NCCL interface exchange, needs comm_init) or
        'halo' (own + halo cells, no communication); the local K / f then hold the final owned columns / dofs."""
        a = self._asm
        # one call = start_assemble + the cell loop + exchange: the local K / f are zero-filled every time unless the bound
        # assembler was created with fillzero=False
        opts = L.AsmOpts(0 if getattr(a, "_accumulate", False) else 1,
                         L.SCATTER_COLORED if a.scatter == "colored" else L.SCATTER_ATOMIC, a.variant, 0)
        L.call("fb2_assemble_distributed", a._handle(self._cv), self.h,
               {"halo": L.DIST_HALO, "exchange": L.DIST_EXCHANGE, "own": L.DIST_OWN_ONLY}[mode],
               element.elem_id, C.byref(element.params), C.sizeof(element.params),
               C.c_void_p(u.data_ptr()) if u is not None else None, C.c_void_p(a.K.nzval.data_ptr()),
               C.c_void_p(a.f.data_ptr()) if a.f is not None else None, C.byref(opts))

    def owned_triplets(self, K, f=None):
        """Owned columns of the local matrix as global (row, col, value) triplets (1-based) + owned f entries."""
        K.dh.grid.ctx.synchronize()
        colptr, rowval = K.colptr - 1, K.rowval - 1
        nz = K.nzval.cpu().numpy()
        cols = np.repeat(np.arange(K.n), np.diff(colptr))
        sel = self.dof_owner[cols] == self.rank
        out = (self.l2g_dof[rowval[sel]], self.l2g_dof[cols[sel]], nz[sel])
        if f is None:
            return out
        own = self.dof_owner == self.rank
        return out + (self.l2g_dof[own], f.cpu().numpy()[own])

    def __del__(self):
        if _destroy is not None:      # module globals are already gone at interpreter shutdown
            _destroy(self, "fb2_partition_destroy", list(_chain(self, "gdh")) + ([self._hctx] if getattr(self, "_hctx", None) is not None else []))


def comm_unique_id():
    buf = (C.c_char * 128)()
    L.call("fb2_comm_unique_id", buf)
    return bytes(buf)


def comm_init(ctx, unique_id, nranks, rank):
    """ncclCommInitRank for the library's own transport; `unique_id` = bytes from comm_unique_id() on rank 0."""
    buf = (C.c_char * 128).from_buffer_copy(unique_id)
    L.call("fb2_comm_init_rank", ctx.h, buf, int(nranks), int(rank))
