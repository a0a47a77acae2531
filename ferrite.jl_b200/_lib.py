"""ctypes bindings of libferrite_b200.so (the C ABI in include/ferrite_b200.h).

The shared library is the product; this module only declares its prototypes.  There is no
CPU fallback: if the library is missing, importing fails loudly; if no CUDA device is usable,
every compute entry point returns FB2_ERR_CUDA which is raised as FB2Error.
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libferrite_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ferrite_b200.h")


class FB2Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[fb2 status {code}] {msg}")
        self.code = code
        self.msg = msg


class DetJNotPositive(FB2Error):
    """throw_detJ_not_pos, src/FEValues/common_values.jl:5"""


class MissingPatternEntry(FB2Error):
    """_missing_sparsity_pattern_error, src/assembler.jl:459-467"""


OK, ERR_BAD_ARG, ERR_CUDA, ERR_OOM, ERR_DETJ, ERR_MISSING, ERR_UNSUPPORTED, ERR_NCCL, ERR_INTERNAL = range(9)
LINE, TRIANGLE, QUADRILATERAL, TETRAHEDRON, HEXAHEDRON = 1, 2, 3, 4, 5
ELEM_HEAT, ELEM_MASS, ELEM_ELASTICITY, ELEM_NEOHOOKE, ELEM_ELASTICITY_GENERAL = 1, 2, 3, 4, 5
SCATTER_ATOMIC, SCATTER_COLORED = 0, 1
BC_FACET, BC_FACE, BC_EDGE, BC_VERTEX, BC_NODE = 0, 1, 2, 3, 4


class Field(C.Structure):
    _fields_ = [("order", C.c_int), ("vdim", C.c_int)]


class HeatParams(C.Structure):
    _fields_ = [("k", C.c_double), ("source", C.c_double)]


class MassParams(C.Structure):
    _fields_ = [("rho", C.c_double)]


class ElasticityParams(C.Structure):
    _fields_ = [("lam", C.c_double), ("mu", C.c_double), ("b", C.c_double * 3)]


class ElasticityGeneralParams(C.Structure):
    _fields_ = [("C", C.c_double * 81), ("b", C.c_double * 3)]


class AsmOpts(C.Structure):
    _fields_ = [("fillzero", C.c_int), ("scatter_mode", C.c_int), ("variant", C.c_int), ("reserved", C.c_int)]


def declared_symbols():
    """Every function name declared in include/ferrite_b200.h."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fb2_[a-z0-9_]+)\s*\(", src)))


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C ferrite.jl_b200/csrc`). There is no CPU fallback.")
    return C.CDLL(LIB_PATH)


lib = load()

_p = C.c_void_p
_pp = C.POINTER(C.c_void_p)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_ip = C.POINTER(C.c_int)
_dp = C.POINTER(C.c_double)

_PROTOS = {
    "fb2_ctx_create": [C.c_int, _pp],
    "fb2_ctx_destroy": [_p],
    "fb2_ctx_synchronize": [_p],
    "fb2_ctx_set_stream": [_p, _p],
    "fb2_ctx_launch_count": [_p, _i64p],
    "fb2_device_alloc": [_p, C.c_size_t, _pp],
    "fb2_device_free": [_p, _p],
    "fb2_memcpy_h2d": [_p, _p, _p, C.c_size_t],
    "fb2_memcpy_d2h": [_p, _p, _p, C.c_size_t],
    "fb2_grid_from_host": [_p, C.c_int, C.c_int64, C.c_int64, C.c_int, _i64p, _dp, _pp],
    "fb2_grid_generate": [_p, C.c_int, _i64p, _dp, _dp, _pp],
    "fb2_grid_perturb": [_p, C.c_double],
    "fb2_grid_set_coordinates": [_p, _dp],
    "fb2_grid_upload_coordinates_async": [_p, _p],
    "fb2_measure_fp64_peak": [_p, _dp],
    "fb2_grid_info": [_p, _ip, _i64p, _i64p, _ip, _ip],
    "fb2_grid_export": [_p, _i64p, _dp],
    "fb2_grid_facetset": [_p, C.c_char_p, _i64p, _i64p],
    "fb2_grid_destroy": [_p],
    "fb2_dh_close": [_p, C.c_int, C.POINTER(Field), _pp],
    "fb2_dh_from_host": [_p, C.c_int, C.POINTER(Field), C.c_int64, C.c_int, _i64p, _pp],
    "fb2_dh_renumber": [_p, C.c_int, _i64p, C.c_int, _i64p, _i64p],
    "fb2_ch_renumber": [_p, _i64p],
    "fb2_dh_info": [_p, _i64p, _ip, _ip],
    "fb2_dh_export": [_p, _i64p],
    "fb2_dh_dof_range": [_p, C.c_int, _ip, _ip],
    "fb2_dh_destroy": [_p],
    "fb2_pattern_create": [_p, _pp],
    "fb2_pattern_from_host": [_p, _i64p, _i64p, _pp],
    "fb2_pattern_info": [_p, _i64p, _i64p],
    "fb2_pattern_export": [_p, _i64p, _i64p],
    "fb2_pattern_destroy": [_p],
    "fb2_cellvalues_create": [_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _pp],
    "fb2_cellvalues_from_tables": [_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _pp],
    "fb2_cellvalues_info": [_p, _ip, _ip, _ip, _ip, _ip],
    "fb2_cellvalues_export": [_p, _dp, _dp, _dp, _dp, _dp, _dp],
    "fb2_cellvalues_destroy": [_p],
    "fb2_function_values": [_p, _p, _p, _p, _p],
    "fb2_reinit_cells": [_p, _p, _i64p, C.c_int64, _p, _p],
    "fb2_spatial_coordinates": [_p, _p, _i64p, C.c_int64, _p],
    "fb2_facetvalues_create": [_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _pp],
    "fb2_facetvalues_info": [_p, _ip, _ip, _ip, _ip, _ip],
    "fb2_facetvalues_export": [_p, _dp, _dp, _dp],
    "fb2_facetvalues_destroy": [_p],
    "fb2_facetset_create": [_p, _i64p, C.c_int64, _pp],
    "fb2_facetset_destroy": [_p],
    "fb2_assemble_facets": [_p, _p, _p, C.c_int, _dp, C.c_int, _p],
    "fb2_spmv": [_p, _p, _p, _p, C.c_int],
    "fb2_csr_values": [_p, _p, _p],
    "fb2_cg": [_p, _p, _p, _p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _ip, _dp],
    "fb2_assembler_create": [_p, _p, _p, _pp],
    "fb2_assemble": [_p, C.c_int, _p, C.c_size_t, _p, _p, _p, C.POINTER(AsmOpts)],
    "fb2_assemble_host": [_p, C.c_int, _p, C.c_size_t, _dp, _dp, _dp, C.POINTER(AsmOpts)],
    "fb2_assemble_host_streamed": [_p, C.c_int, _p, C.c_size_t, _dp, _dp, _dp, _dp, C.POINTER(AsmOpts)],
    "fb2_assembler_coloring": [_p, _ip, _i32p],
    "fb2_scatter_host": [_p, _dp, _dp, _p, _p, C.POINTER(AsmOpts)],
    "fb2_scatter_device": [_p, _p, _p, _p, _p, C.POINTER(AsmOpts)],
    "fb2_ea_create": [_p, _p, _pp],
    "fb2_ea_info": [_p, _i64p, _ip],
    "fb2_ea_assemble": [_p, C.c_int, _p, C.c_size_t, _p, _p, _p],
    "fb2_ea_mul": [_p, _p, _p, _p],
    "fb2_ea_apply_local": [_p, _p, _p, _p, C.c_int],
    "fb2_apply_assemble": [_p, _p, _p, C.c_int, _p, C.c_size_t, _p, _p, _p, C.c_int, C.POINTER(AsmOpts)],
    "fb2_ea_diag": [_p, _p, _p],
    "fb2_ea_rhs": [_p, _p, _p],
    "fb2_ea_cg": [_p, _p, _p, _p, C.c_double, C.c_double, C.c_int, C.c_int, _ip, _dp],
    "fb2_ea_destroy": [_p],
    "fb2_rhsdata_create": [_p, _p, _p, _pp],
    "fb2_rhsdata_info": [_p, _dp, _i64p, _i64p],
    "fb2_apply_rhs": [_p, _p, _p, C.c_int],
    "fb2_rhsdata_destroy": [_p],
    "fb2_assembler_destroy": [_p],
    "fb2_ch_create": [_p, _pp],
    "fb2_ch_add_dirichlet": [_p, C.c_int, C.c_int, C.c_int64, _i64p, C.c_int, _ip, _ip],
    "fb2_ch_close": [_p],
    "fb2_ch_from_host": [_p, C.c_int64, _i64p, _dp, _pp],
    "fb2_ch_set_inhomogeneities": [_p, C.c_int64, _dp],
    "fb2_ch_bc_points": [_p, C.c_int, _i64p, _dp],
    "fb2_ch_bc_set_values": [_p, C.c_int, C.c_int64, _dp],
    "fb2_assemble_mixed_up": [_p, _p, _p, C.c_int, C.c_int, C.c_double, C.c_double, _p, _p, C.POINTER(AsmOpts)],
    "fb2_ch_add_affine": [_p, C.c_int64, C.c_int, _i64p, _dp, C.c_double],
    "fb2_ch_add_periodic": [_p, C.c_int, C.c_int64, _i64p, C.c_int64, _i64p, C.c_int, _ip],
    "fb2_ch_affine_export": [_p, _i64p, _i64p, _i64p, _dp],
    "fb2_pattern_create_condensed": [_p, _p, _pp],
    "fb2_ch_info": [_p, _i64p],
    "fb2_ch_export": [_p, _i64p, _dp],
    "fb2_apply": [_p, _p, _p, _p, C.c_int, _dp],
    "fb2_apply_vector": [_p, _p, C.c_int],
    "fb2_ch_destroy": [_p],
    "fb2_partition_create": [_p, C.c_int, C.c_int, _ip, _pp],
    "fb2_partition_create_from_owners": [_p, C.c_int, C.c_int, _i32p, _pp],
    "fb2_partition_create_metis": [_p, C.c_int, C.c_int, _pp],
    "fb2_partition_create_generated": [_p, _i64p, _dp, _dp, C.c_double, C.c_int, C.c_int, C.c_int, _ip, _pp],
    "fb2_partition_info": [_p, _i64p, _i64p, _i64p, _i64p, _i64p],
    "fb2_partition_export": [_p, _i64p, C.POINTER(C.c_uint8), _i64p, _i64p, _i32p],
    "fb2_partition_local_grid": [_p, _p, _pp],
    "fb2_partition_local_dh": [_p, _p, _pp],
    "fb2_partition_peer_counts": [_p, C.c_int, _i64p, _i64p, _i64p, _i64p],
    "fb2_partition_peer_lists": [_p, C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p],
    "fb2_partition_bind": [_p, _p],
    "fb2_partition_pack": [_p, C.c_int, _p, _p, _p],
    "fb2_partition_unpack_add": [_p, C.c_int, _p, _p, _p],
    "fb2_partition_mask_unowned": [_p, _p, _p],
    "fb2_partition_destroy": [_p],
    "fb2_comm_unique_id": [_p],
    "fb2_comm_init_rank": [_p, _p, C.c_int, C.c_int],
    "fb2_comm_destroy": [_p],
    "fb2_partition_exchange": [_p, _p, _p],
    "fb2_assemble_distributed": [_p, _p, C.c_int, C.c_int, _p, C.c_size_t, _p, _p, _p, C.POINTER(AsmOpts)],
}
DIST_EXCHANGE, DIST_HALO, DIST_OWN_ONLY = 0, 1, 2
FACET_FLUX, FACET_TRACTION, FACET_NORMAL_TRACTION = 1, 2, 3
ORDER_PERMUTATION, ORDER_FIELDWISE, ORDER_COMPONENTWISE, ORDER_METIS = 0, 1, 2, 3

lib.fb2_version.restype = C.c_char_p
lib.fb2_version.argtypes = []
lib.fb2_last_error.restype = C.c_char_p
lib.fb2_last_error.argtypes = []
lib.fb2_last_kernel.restype = C.c_char_p
lib.fb2_last_kernel.argtypes = []
for _name, _args in _PROTOS.items():
    _f = getattr(lib, _name)
    _f.argtypes = _args
    _f.restype = C.c_int


def check(rc):
    if rc == OK:
        return
    msg = lib.fb2_last_error().decode("utf-8", "replace")
    if rc == ERR_DETJ:
        raise DetJNotPositive(rc, msg)
    if rc == ERR_MISSING:
        raise MissingPatternEntry(rc, msg)
    raise FB2Error(rc, msg)


def call(name, *args):
    check(getattr(lib, name)(*args))
