# FerriteB200.jl -- thin Julia shim over libferrite_b200.so (the C ABI in include/ferrite_b200.h).
#
# NOT EXECUTED in this repository's CI: neither the build container nor the GPU box has a Julia toolchain.
# It is the binding a Ferrite.jl maintainer would add; it overloads the reference's documented extension points
# (docs/src/devdocs/assembly.md:18-51): allocate_matrix, start_assemble, assemble!, finish_assemble, apply!.
# The executed mirror of the same calls is ferrite.jl_b200/api.py (ctypes); tests/ drive the library through it.
module FerriteB200

using Ferrite
using SparseArrays

const LIB = get(ENV, "FERRITE_B200_LIB", "libferrite_b200")

struct FB2Error <: Exception
    code::Cint
    msg::String
end

function check(rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:fb2_last_error, LIB), Cstring, ()))
    rc == 4 && throw(ArgumentError(msg))      # FB2_ERR_DETJ_NOT_POSITIVE <-> throw_detJ_not_pos (common_values.jl:5)
    rc == 5 && throw(ErrorException(msg))     # FB2_ERR_MISSING_PATTERN_ENTRY <-> assembler.jl:459-467
    throw(FB2Error(rc, msg))
end

macro fb2(f, argtypes, args...)
    return :(check(ccall(($(QuoteNode(f)), LIB), Cint, $(esc(argtypes)), $(map(esc, args)...))))
end

mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = 0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        @fb2 fb2_ctx_create (Cint, Ptr{Ptr{Cvoid}}) device r
        return finalizer(c -> ccall((:fb2_ctx_destroy, LIB), Cint, (Ptr{Cvoid},), c.h), new(r[]))
    end
end

const CELLTYPE = Dict(Line => 1, Triangle => 2, Quadrilateral => 3, Tetrahedron => 4, Hexahedron => 5)

struct fb2_field
    order::Cint
    vdim::Cint
end

# ---- arrays-in mode: keep Ferrite for set-up, hand its arrays to the device ----------------------------------------
"""
    DeviceProblem(ctx, dh, cv)

Uploads `dh.grid`, `dh.cell_dofs` (src/Dofs/DofHandler.jl:126-131) and the tables of `cv`
(cv.fun_values.Nξ/dNdξ FunctionValues.jl:46-49, cv.geo_mapping.M/dMdξ GeometryMapping.jl:38-39, cv.qr.weights).
For vector interpolations pass the CellValues of the scalar base interpolation; `vdim` is taken from `dh`.
"""
mutable struct DeviceProblem
    ctx::Context
    grid::Ptr{Cvoid}
    dh::Ptr{Cvoid}
    cv::Ptr{Cvoid}
    ndofs::Int
end

function DeviceProblem(ctx::Context, dh::DofHandler{sdim}, cv::CellValues, vdim::Integer = 1) where {sdim}
    grid = dh.grid
    CT = getcelltype(grid)
    cells = reinterpret(reshape, Int, grid.cells)             # nnpc x ncells, 1-based (Vector{Hexahedron} is isbits)
    xyz = reinterpret(reshape, Float64, grid.nodes)           # sdim x nnodes
    g = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve cells xyz begin
        @fb2 fb2_grid_from_host (Ptr{Cvoid}, Cint, Int64, Int64, Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Ptr{Cvoid}}) ctx.h CELLTYPE[CT] getncells(grid) getnnodes(grid) sdim cells xyz g
    end
    ip = Ferrite.getfieldinterpolation(dh.subdofhandlers[1], 1)
    fields = [fb2_field(Ferrite.getorder(ip), vdim)]
    d = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve fields begin
        @fb2 fb2_dh_from_host (Ptr{Cvoid}, Cint, Ptr{fb2_field}, Int64, Cint, Ptr{Int64}, Ptr{Ptr{Cvoid}}) g[] 1 fields ndofs(dh) ndofs_per_cell(dh) dh.cell_dofs d
    end
    N = cv.fun_values.Nξ                                       # n x nq
    dN = reinterpret(reshape, Float64, cv.fun_values.dNdξ)     # rdim x n x nq
    M = cv.geo_mapping.M
    dM = reinterpret(reshape, Float64, cv.geo_mapping.dMdξ)
    w = cv.qr.weights
    c = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve N dN M dM w begin
        @fb2 fb2_cellvalues_from_tables (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Ptr{Cvoid}}) ctx.h CELLTYPE[CT] length(w) size(N, 1) vdim size(M, 1) N dN M dM w c
    end
    return DeviceProblem(ctx, g[], d[], c[], ndofs(dh))
end

# ---- matrix type: plugs into allocate_matrix / start_assemble / assemble! / apply! -----------------------------------
"""
    B200Matrix <: AbstractSparseMatrix

`K.host::SparseMatrixCSC{Float64,Int}` carries colptr/rowval (bit-identical to `allocate_matrix(dh)`); the values
live on the device until `finish_assemble` / `download!` copies them into `K.host.nzval`.
"""
mutable struct B200Matrix
    prob::DeviceProblem
    pattern::Ptr{Cvoid}
    assembler::Ptr{Cvoid}
    host::SparseMatrixCSC{Float64, Int}
    f::Vector{Float64}
end

# allocate_matrix(::Type{B200Matrix}, dh) -- src/Dofs/sparsity_pattern.jl:628-645
function Ferrite.allocate_matrix(::Type{B200Matrix}, prob::DeviceProblem)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    @fb2 fb2_pattern_create (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}) prob.dh p          # pattern built on the device
    n = Ref{Int64}(0); nnz = Ref{Int64}(0)
    @fb2 fb2_pattern_info (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}) p[] n nnz
    colptr = Vector{Int}(undef, n[] + 1); rowval = Vector{Int}(undef, nnz[])
    @fb2 fb2_pattern_export (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}) p[] colptr rowval
    a = Ref{Ptr{Cvoid}}(C_NULL)
    @fb2 fb2_assembler_create (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}) prob.dh p[] prob.cv a
    return B200Matrix(prob, p[], a[], SparseMatrixCSC(n[], n[], colptr, rowval, zeros(nnz[])), zeros(n[]))
end

# element menu (the reference's element routine is user code; see DESIGN.md section 1)
struct HeatElement;       k::Float64; source::Float64; end
struct ElasticityElement; lambda::Float64; mu::Float64; b::NTuple{3, Float64}; end
elem_id(::HeatElement) = Cint(1)
elem_id(::ElasticityElement) = Cint(3)

struct fb2_asm_opts
    fillzero::Cint
    scatter_mode::Cint
    variant::Cint
    reserved::Cint
end

# `fillzero` is a pending flag: start_assemble zeroes K and f ONCE (src/assembler.jl:287-291); the fill is fused into the
# first assemble! on this assembler and cleared, later assemble! calls add onto the result like the reference's.
mutable struct B200Assembler <: Ferrite.AbstractAssembler
    K::B200Matrix
    fillzero::Bool
end
take_fillzero!(a::B200Assembler) = (z = a.fillzero; a.fillzero = false; z)

# start_assemble(K, f; fillzero) -- src/assembler.jl:287-291
Ferrite.start_assemble(K::B200Matrix; fillzero::Bool = true) = B200Assembler(K, fillzero)

"""
    assemble!(assembler, element)

The whole `for cell in CellIterator(dh) ... assemble!(assembler, celldofs(cell), Ke, fe)` loop
(heat_equation.jl:181-204) in one call; results land in `K.host.nzval` and `K.f`.
"""
function Ferrite.assemble!(a::B200Assembler, element; u::Union{Nothing, Vector{Float64}} = nothing)
    K = a.K
    opts = Ref(fb2_asm_opts(take_fillzero!(a), 0, 0, 0))
    params = Ref(element)
    GC.@preserve params opts begin
        @fb2 fb2_assemble_host (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Csize_t, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{fb2_asm_opts}) K.assembler elem_id(element) params sizeof(element) (u === nothing ? C_NULL : pointer(u)) K.host.nzval K.f opts
    end
    return a
end

Ferrite.finish_assemble(a::B200Assembler) = (a.K.host, a.K.f)

# apply!(K, f, ch) -- src/Dofs/ConstraintHandler.jl:710-740; the closed reference ConstraintHandler is adopted as arrays
function Ferrite.apply!(K::B200Matrix, ch::ConstraintHandler, nzval_dev::Ptr{Float64}, f_dev::Ptr{Float64}; applyzero::Bool = false)
    c = Ref{Ptr{Cvoid}}(C_NULL)
    @fb2 fb2_ch_from_host (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Ptr{Cvoid}}) K.prob.dh length(ch.prescribed_dofs) ch.prescribed_dofs ch.inhomogeneities c
    m = Ref{Float64}(0.0)
    @fb2 fb2_apply (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}) c[] K.pattern nzval_dev f_dev applyzero m
    ccall((:fb2_ch_destroy, LIB), Cint, (Ptr{Cvoid},), c[])
    return m[]
end

# A closed reference ConstraintHandler WITH affine / periodic constraints (ch.dofcoefficients, src/Dofs/ConstraintHandler.jl:160-165):
# rebuilt natively from its own fields -- Dirichlet dofs as value-only constraints, affine ones with their masters -- so that
# apply! runs `_condense!` on the device and allocate_matrix(dh, ch) returns the condensed pattern.
function native_constraints(dh_native::Ptr{Cvoid}, ch::ConstraintHandler)
    c = Ref{Ptr{Cvoid}}(C_NULL)
    @fb2 fb2_ch_create (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}) dh_native c
    for (i, dof) in pairs(ch.prescribed_dofs)
        coeffs = ch.dofcoefficients[i]
        masters = coeffs === nothing ? Int64[] : Int64[first(p) for p in coeffs]
        vals = coeffs === nothing ? Float64[] : Float64[last(p) for p in coeffs]
        @fb2 fb2_ch_add_affine (Ptr{Cvoid}, Int64, Cint, Ptr{Int64}, Ptr{Float64}, Float64) c[] dof length(masters) masters vals ch.inhomogeneities[i]
    end
    @fb2 fb2_ch_close (Ptr{Cvoid},) c[]
    return c[]
end

# allocate_matrix(dh, ch) -- src/Dofs/sparsity_pattern.jl:628-645 with `_add_constraint_entries!` :782-844
function condensed_pattern(dh_native::Ptr{Cvoid}, ch_native::Ptr{Cvoid})
    p = Ref{Ptr{Cvoid}}(C_NULL)
    @fb2 fb2_pattern_create_condensed (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}) dh_native ch_native p
    return p[]
end

# get_rhs_data(ch, A) / apply_rhs!(data, f, ch, applyzero) -- src/Dofs/ConstraintHandler.jl:191-240 (one factorisation, many steps)
mutable struct B200RHSData
    h::Ptr{Cvoid}
    ch::Ptr{Cvoid}
end

function Ferrite.get_rhs_data(ch::ConstraintHandler, K::B200Matrix, nzval_dev::Ptr{Float64})
    c = Ref{Ptr{Cvoid}}(C_NULL)
    @fb2 fb2_ch_from_host (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Ptr{Cvoid}}) K.prob.dh length(ch.prescribed_dofs) ch.prescribed_dofs ch.inhomogeneities c
    r = Ref{Ptr{Cvoid}}(C_NULL)
    @fb2 fb2_rhsdata_create (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Ptr{Cvoid}}) c[] K.pattern nzval_dev r
    return finalizer(B200RHSData(r[], c[])) do d
        ccall((:fb2_rhsdata_destroy, LIB), Cint, (Ptr{Cvoid},), d.h)
        ccall((:fb2_ch_destroy, LIB), Cint, (Ptr{Cvoid},), d.ch)
    end
end

function Ferrite.apply_rhs!(data::B200RHSData, f_dev::Ptr{Float64}, ch::ConstraintHandler, applyzero::Bool = false)
    # the inhomogeneities of the current update!(ch, t) travel as one small array
    @fb2 fb2_ch_set_inhomogeneities (Ptr{Cvoid}, Int64, Ptr{Float64}) data.ch length(ch.inhomogeneities) ch.inhomogeneities
    @fb2 fb2_apply_rhs (Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}, Cint) data.h f_dev data.ch applyzero
    return f_dev
end

# ---- Neumann / traction facet loop (hyperelasticity.jl:278-291) ------------------------------------------------------
# `facetset` is a reference facet set (OrderedSet{FacetIndex}); kind: 1 flux (params q), 2 traction (params t),
# 3 normal traction (params p: fe += p n N dGamma; the tutorial's `ge[i] -= (dui . tn n) dGamma` is p = -tn).
# f_dev is the device residual / load vector the volume assembly wrote.
function assemble_facets!(f_dev::Ptr{Float64}, prob::DeviceProblem, celltype::Integer, ip_order::Int, vdim::Int, qr_order::Int,
        facetset, kind::Integer, params::Vector{Float64})
    pairs = Matrix{Int64}(undef, 2, length(facetset))
    for (k, fi) in enumerate(facetset)
        pairs[1, k], pairs[2, k] = fi[1], fi[2]
    end
    fv = Ref{Ptr{Cvoid}}(C_NULL)
    set = Ref{Ptr{Cvoid}}(C_NULL)
    @fb2 fb2_facetvalues_create (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint, Ptr{Ptr{Cvoid}}) prob.ctx.h celltype qr_order ip_order vdim 1 fv
    @fb2 fb2_facetset_create (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Ptr{Cvoid}}) prob.grid pairs size(pairs, 2) set
    @fb2 fb2_assemble_facets (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Ptr{Float64}) prob.dh fv[] set[] kind params length(params) f_dev
    ccall((:fb2_facetset_destroy, LIB), Cint, (Ptr{Cvoid},), set[])
    ccall((:fb2_facetvalues_destroy, LIB), Cint, (Ptr{Cvoid},), fv[])
    return f_dev
end

# ---- the consumer of K on the device: IterativeSolvers.cg!(x, K, b; ...) (hyperelasticity.jl:418), mul! ------------------
# nzval_dev / b_dev / x_dev are device pointers (what fb2_assemble wrote); returns (iterations, final residual norm)
function cg_device!(x_dev::Ptr{Float64}, K::B200Matrix, nzval_dev::Ptr{Float64}, b_dev::Ptr{Float64};
        reltol::Float64 = sqrt(eps(Float64)), abstol::Float64 = 0.0, maxiter::Int = size(K.host, 1), jacobi::Bool = false)
    iters = Ref{Cint}(0)
    res = Ref{Float64}(0.0)
    @fb2 fb2_cg (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cdouble, Cdouble, Cint, Cint, Cint, Ptr{Cint}, Ptr{Float64}) K.pattern nzval_dev b_dev x_dev reltol abstol maxiter jacobi true iters res
    return Int(iters[]), res[]
end

function mul_device!(y_dev::Ptr{Float64}, K::B200Matrix, nzval_dev::Ptr{Float64}, x_dev::Ptr{Float64}; transpose::Bool = false)
    @fb2 fb2_spmv (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint) K.pattern nzval_dev x_dev y_dev transpose
    return y_dev
end

# ---- element assembly (docs/src/literate-howto/gpu_assembly.jl:265-304) and apply_assemble! (src/assembler.jl:491-503) -----
# Kes_dev (n x n x ncells) / fes_dev (n x ncells) are device buffers of the caller, e.g. CUDA.jl arrays passed as pointers.
mutable struct ElementAssembly
    h::Ptr{Cvoid}
    prob::DeviceProblem
    function ElementAssembly(prob::DeviceProblem)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        @fb2 fb2_ea_create (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}) prob.dh prob.cv r
        return finalizer(e -> ccall((:fb2_ea_destroy, LIB), Cint, (Ptr{Cvoid},), e.h), new(r[], prob))
    end
end

function element_matrices!(Kes_dev::Ptr{Float64}, fes_dev::Ptr{Float64}, ea::ElementAssembly, element; u_dev::Ptr{Float64} = Ptr{Float64}(C_NULL))
    params = Ref(element)
    GC.@preserve params begin
        @fb2 fb2_ea_assemble (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Csize_t, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}) ea.h elem_id(element) params sizeof(element) u_dev Kes_dev fes_dev
    end
    return Kes_dev, fes_dev
end

# mul!(y, A, x) of the matrix-free operator (gpu_assembly.jl:287-304)
function mul_device!(y_dev::Ptr{Float64}, ea::ElementAssembly, Kes_dev::Ptr{Float64}, x_dev::Ptr{Float64})
    @fb2 fb2_ea_mul (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}) ea.h Kes_dev x_dev y_dev
    return y_dev
end

# f = sum_e P_e' fe and the diagonal of the operator
function rhs_device!(f_dev::Ptr{Float64}, ea::ElementAssembly, fes_dev::Ptr{Float64})
    @fb2 fb2_ea_rhs (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}) ea.h fes_dev f_dev
    return f_dev
end

function diag_device!(d_dev::Ptr{Float64}, ea::ElementAssembly, Kes_dev::Ptr{Float64})
    @fb2 fb2_ea_diag (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}) ea.h Kes_dev d_dev
    return d_dev
end

# IterativeSolvers.cg!(x, A, b) with the matrix-free operator; returns (iterations, final residual norm)
function cg_device!(x_dev::Ptr{Float64}, ea::ElementAssembly, Kes_dev::Ptr{Float64}, b_dev::Ptr{Float64};
        reltol::Float64 = sqrt(eps(Float64)), abstol::Float64 = 0.0, maxiter::Int = ea.prob.ndofs, jacobi::Bool = false)
    iters = Ref{Cint}(0)
    res = Ref{Float64}(0.0)
    @fb2 fb2_ea_cg (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cdouble, Cdouble, Cint, Cint, Ptr{Cint}, Ptr{Float64}) ea.h Kes_dev b_dev x_dev reltol abstol maxiter jacobi iters res
    return Int(iters[]), res[]
end

# apply_local!(Ke, fe, celldofs(cell), ch; apply_zero) for every cell (src/Dofs/ConstraintHandler.jl:1750-1822)
function apply_local_device!(Kes_dev::Ptr{Float64}, fes_dev::Ptr{Float64}, ea::ElementAssembly, ch::ConstraintHandler; apply_zero::Bool = false)
    c = Ref{Ptr{Cvoid}}(C_NULL)
    @fb2 fb2_ch_from_host (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Ptr{Cvoid}}) ea.prob.dh length(ch.prescribed_dofs) ch.prescribed_dofs ch.inhomogeneities c
    @fb2 fb2_ea_apply_local (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint) ea.h c[] Kes_dev fes_dev apply_zero
    ccall((:fb2_ch_destroy, LIB), Cint, (Ptr{Cvoid},), c[])
    return Kes_dev, fes_dev
end

# assemble!(assembler, celldofs(cell), Ke, fe) for every cell from the stored element matrices
function scatter_device!(K::B200Matrix, nzval_dev::Ptr{Float64}, f_dev::Ptr{Float64}, Kes_dev::Ptr{Float64}, fes_dev::Ptr{Float64}; fillzero::Bool = true)
    opts = Ref(fb2_asm_opts(fillzero, 0, 0, 0))
    GC.@preserve opts begin
        @fb2 fb2_scatter_device (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{fb2_asm_opts}) K.assembler Kes_dev fes_dev nzval_dev f_dev opts
    end
    return nzval_dev, f_dev
end

# the cell loop with apply_assemble!(assembler, ch, celldofs(cell), Ke, fe; apply_zero) in place of assemble!
function Ferrite.apply_assemble!(a::B200Assembler, ea::ElementAssembly, ch::ConstraintHandler, element, nzval_dev::Ptr{Float64}, f_dev::Ptr{Float64};
        u_dev::Ptr{Float64} = Ptr{Float64}(C_NULL), apply_zero::Bool = false)
    c = Ref{Ptr{Cvoid}}(C_NULL)
    @fb2 fb2_ch_from_host (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Ptr{Ptr{Cvoid}}) ea.prob.dh length(ch.prescribed_dofs) ch.prescribed_dofs ch.inhomogeneities c
    opts = Ref(fb2_asm_opts(take_fillzero!(a), 0, 0, 0))
    params = Ref(element)
    GC.@preserve params opts begin
        @fb2 fb2_apply_assemble (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Csize_t, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{fb2_asm_opts}) a.K.assembler ea.h c[] elem_id(element) params sizeof(element) u_dev nzval_dev f_dev apply_zero opts
    end
    ccall((:fb2_ch_destroy, LIB), Cint, (Ptr{Cvoid},), c[])
    return a
end

# ---- partition plan of rank `rank` of `nparts` (one process per GPU): METIS_PartMeshDual inside the library, or the
# cell -> rank vector of any partitioner (0-based ranks) -----------------------------------------------------------
function partition_plan(global_dh_handle::Ptr{Cvoid}, nparts::Integer, rank::Integer; cell_owner::Union{Nothing, Vector{Int32}} = nothing)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    if cell_owner === nothing
        @fb2 fb2_partition_create_metis (Ptr{Cvoid}, Cint, Cint, Ptr{Ptr{Cvoid}}) global_dh_handle nparts rank p
    else
        @fb2 fb2_partition_create_from_owners (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Ptr{Ptr{Cvoid}}) global_dh_handle nparts rank cell_owner p
    end
    return p[]
end

# the same plan for generate_grid(Hexahedron, nel, left, right) [+ perturbation] with one Lagrange{RefHexahedron,1}()^vdim field,
# built rank-locally (no global grid / DofHandler): px, py, pz blocks
function partition_plan_generated(host_ctx::Ptr{Cvoid}, nel::NTuple{3, Int}, left::Vec{3, Float64}, right::Vec{3, Float64}, vdim::Integer,
        nparts::Integer, rank::Integer, dims::NTuple{3, Int}; perturb::Float64 = 0.0)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    n, l, r, d = Int64[nel...], Float64[left...], Float64[right...], Cint[dims...]
    @fb2 fb2_partition_create_generated (Ptr{Cvoid}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Cdouble, Cint, Cint, Cint, Ptr{Cint}, Ptr{Ptr{Cvoid}}) host_ctx n l r perturb vdim nparts rank d p
    return p[]
end

end # module