# GRMPCuda.jl -- Julia glue that puts libgrmp_cuda behind the unchanged
# GradientRobustMultiPhysics.jl API (PDEDescription / add_operator! / assemble! / solve!).
#
# NOT RUNNABLE IN THE BUILD CONTAINER (no julia binary, no network).  It is kept mechanical: every `ccall`
# below has a line-for-line ctypes twin in gradientrobustmultiphysics.jl_b200/_lib.py + assembly.py, which is
# what the test-suite runs.
#
# How it hooks in.  The reference assembles through
#     assemble!(A::AbstractArray{T,2}, AP::AssemblyPattern{<:APT_BilinearForm,T,AT,Tv,Ti}, FEB = []; ...)   bilinearform.jl:92-103
#     assemble!(b::Union{AbstractArray{T,1},AbstractArray{T,2}}, AP::AssemblyPattern{<:APT_LinearForm,...}, FEB = []; ...)   linearform.jl:47-54
# and `assemble_operator!` (pdeoperators.jl:978-1006) calls them WITHOUT the FEB argument for every operator that has no
# fixed arguments.  This file adds the two-argument methods
#     assemble!(A::FEMatrixBlock{Float64,Int64,Float64,Int32}, AP::AssemblyPattern{<:APT_BilinearForm,Float64,AT,Float64,Int32}; ...)
#     assemble!(b::FEVectorBlock{Float64,Float64,Int32},       AP::AssemblyPattern{<:APT_LinearForm,Float64,AT,Float64,Int32}; ...)
#   with AT = ON_CELLS, or ON_BFACES for the Identity forms of H1P1 / H1P2 (best-approximation boundary data, boundarydata.jl:297-347)
# which are more specific than the reference's, so PDEDescription / add_operator! / solve! reach them unchanged.  Calls WITH
# coefficient arguments (FEB, row N4 of SURVEY.md 8f) dispatch to a third method further down that takes the Picard convection form
# of ConvectionOperator and hands everything else back.  Inside, `blf_plan` / `lf_plan` decide whether the
# (FEType, operator, action) triple is on the device path; everything else is handed back to the reference method with
# `invoke` and ALL keyword arguments -- the library itself has no CPU fallback, the reference loop simply stays reachable.
module GRMPCuda

using GradientRobustMultiPhysics
using ExtendableGrids
using ExtendableSparse
using SparseArrays

const GRMP = GradientRobustMultiPhysics
const lib = get(ENV, "LIBGRMP_CUDA", "libgrmp_cuda")

struct GrmpError <: Exception
    code::Cint
    msg::String
end
check(rc::Cint) = rc == 0 ? nothing : throw(GrmpError(rc, unsafe_string(ccall((:grmp_last_error, lib), Cstring, ()))))

# ---- codes of include/grmp.h (nothing = not on the device path) -----------------------------------------------------
fecode(::Type) = nothing
fecode(::Type{<:H1P1}) = 1
fecode(::Type{<:H1P2}) = 2
fecode(::Type{<:H1Pk{n,2,2}}) where {n} = 2     # same tables as H1P2 (DESIGN.md 3.2)
fecode(::Type{<:H1Pk{n,3,2}}) where {n} = 2
fecode(::Type{<:H1BR}) = 3
fecode(::Type{<:HDIVRT0}) = 4
fecode(::Type{<:HDIVBDM1}) = 5
fecode(::Type{<:L2P0}) = 6
opcode(::Type) = nothing
opcode(::Type{Identity}) = 1
opcode(::Type{Gradient}) = 2
opcode(::Type{SymmetricGradient{1}}) = 3
opcode(::Type{Divergence}) = 4
opcode(::Type{ReconstructionIdentity{FER}}) where {FER<:HDIVRT0} = 5
opcode(::Type{ReconstructionIdentity{FER}}) where {FER<:HDIVBDM1} = 6
opcode(::Type{NormalFlux}) = 7                  # Hdiv spaces on boundary faces only (the library refuses it elsewhere)
aptcode(::Type{GRMP.APT_BilinearForm}) = 0
aptcode(::Type{GRMP.APT_SymmetricBilinearForm}) = 1
aptcode(::Type{GRMP.APT_LumpedBilinearForm}) = 2
const F_NONE, F_CONST, F_QP_TABLE = 0, 1, 2

struct EvalTab
    nd_all::Int32
    ncomp::Int32
    refvals::Ptr{Float64}
    refderivs::Ptr{Float64}
end

# ---- handles: the library owns the memory, finalizers call the *_destroy entry points (grmp.h "Ownership") ------------------
mutable struct Ctx;    h::Ptr{Cvoid}; device::Int; end
mutable struct DGrid;  h::Ptr{Cvoid}; hasfaces::Bool; ctx::Ctx; end
mutable struct DSpace; h::Ptr{Cvoid}; grid::DGrid; end
mutable struct DBlf
    h::Ptr{Cvoid}; nnz::Int64; colptr::Vector{Int64}; rowval::Vector{Int64}
    factor::Float64; transposed::Bool; spaces::Tuple{DSpace,DSpace}
end
mutable struct DLf; h::Ptr{Cvoid}; space::DSpace; end
destroy(sym::Symbol, x) = (x.h == C_NULL || ccall((sym, lib), Cint, (Ptr{Cvoid},), x.h); x.h = C_NULL; nothing)

# One context per device of THIS process (grmp_init(device, &ctx)).  A single Julia process reaches several GPUs by making
# one of them current -- `GRMPCuda.device!(k)` -- before the grid of a subdomain is first assembled: every handle remembers
# its context, so forms on different grids may live on different devices and be assembled from different tasks
# (handles are independent, grmp.h "Threading"; ccall blocks only the calling task's thread).  Multi-GPU assembly of ONE grid
# is one process per GPU (DESIGN.md 4): the partition is made on the host, every process runs this same glue on its part.
const CONTEXTS = Dict{Int,Ctx}()
const CURRENT_DEVICE = Ref(0)
device!(k::Integer) = (CURRENT_DEVICE[] = Int(k); context(); nothing)
function context(device::Int = CURRENT_DEVICE[])
    get!(CONTEXTS, device) do
        h = Ref{Ptr{Cvoid}}()
        check(ccall((:grmp_init, lib), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h))
        c = Ctx(h[], device)
        finalizer(x -> destroy(:grmp_finalize, x), c)
        c
    end
end

# caches are weak in the Julia object: a grid / space / pattern that is garbage-collected releases its device memory
const GRIDS = WeakKeyDict{Any,DGrid}()
function device_grid(xgrid::ExtendableGrid{Float64,Int32}; faces = false)
    g = get!(GRIDS, xgrid) do
        coords = xgrid[Coordinates]; cn = Matrix{Int32}(xgrid[CellNodes])
        vol = Vector{Float64}(xgrid[CellVolumes]); reg = Vector{Int32}(xgrid[CellRegions])
        h = Ref{Ptr{Cvoid}}()
        ctx = context()
        GC.@preserve coords cn vol reg check(ccall((:grmp_grid_create, lib), Cint,
            (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}, Ref{Ptr{Cvoid}}),
            ctx.h, size(coords, 1), size(coords, 2), coords, size(cn, 2), cn, vol, reg, h))
        d = DGrid(h[], false, ctx)
        finalizer(x -> destroy(:grmp_grid_destroy, x), d)
        d
    end
    if faces && !g.hasfaces
        cf = Matrix{Int32}(xgrid[CellFaces]); sg = Matrix{Int32}(xgrid[CellFaceSigns])
        ori = size(xgrid[Coordinates], 1) == 3 ? Matrix{Int32}(xgrid[CellFaceOrientations]) : Matrix{Int32}(undef, 0, 0)
        fn = Matrix{Float64}(xgrid[FaceNormals]); fv = Vector{Float64}(xgrid[FaceVolumes])
        GC.@preserve cf sg ori fn fv check(ccall((:grmp_grid_set_faces, lib), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
            g.h, length(fv), cf, sg, isempty(ori) ? C_NULL : pointer(ori), fn, fv))
        g.hasfaces = true
    end
    return g
end

# AT = ON_BFACES: the boundary faces as items (BFaceNodes / BFaceVolumes / BFaceRegions, assemblypatterns.jl:400-440)
const BGRIDS = WeakKeyDict{Any,DGrid}()
function device_grid(xgrid::ExtendableGrid{Float64,Int32}, ::Type{ON_BFACES})
    get!(BGRIDS, xgrid) do
        coords = xgrid[Coordinates]; bn = Matrix{Int32}(xgrid[BFaceNodes])
        vol = Vector{Float64}(xgrid[BFaceVolumes]); reg = Vector{Int32}(xgrid[BFaceRegions])
        h = Ref{Ptr{Cvoid}}()
        ctx = context()
        GC.@preserve coords bn vol reg check(ccall((:grmp_grid_create_bfaces, lib), Cint,
            (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}, Ref{Ptr{Cvoid}}),
            ctx.h, size(coords, 1), size(coords, 2), coords, size(bn, 2), bn, vol, reg, h))
        d = DGrid(h[], false, ctx)
        finalizer(x -> destroy(:grmp_grid_destroy, x), d)
        d
    end
end
device_grid(xgrid::ExtendableGrid{Float64,Int32}, ::Type{ON_CELLS}; faces = false) = device_grid(xgrid; faces)

"`update_geometry!(xgrid)`: the coordinates of a grid moved in place (same topology) -- re-upload Coordinates / CellVolumes"
function update_geometry!(xgrid::ExtendableGrid{Float64,Int32})
    haskey(GRIDS, xgrid) || return nothing
    coords = xgrid[Coordinates]; vol = Vector{Float64}(xgrid[CellVolumes])
    GC.@preserve coords vol check(ccall((:grmp_grid_update_geometry, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), GRIDS[xgrid].h, coords, vol))
end

const SPACES = WeakKeyDict{Any,DSpace}()
function device_space(FES::FESpace{Float64,Int32,FEType}) where {FEType}
    get!(SPACES, FES) do
        g = device_grid(FES.xgrid; faces = FEType <: Union{H1BR,HDIVRT0,HDIVBDM1})
        dofs = FES[CellDofs]
        ncells = num_sources(FES.xgrid[CellNodes])
        # broken spaces (always L2P0, finiteelements.jl:78-80) carry a SerialVariableTargetAdjacency: dofs of cell c are (c-1) nd + 1 : c nd
        colentries = dofs isa GRMP.SerialVariableTargetAdjacency ? Matrix{Int32}(reshape(Int32(1):Int32(FES.ndofs), :, ncells)) :
                     Matrix{Int32}(reshape(dofs.colentries, :, ncells))
        h = Ref{Ptr{Cvoid}}()
        GC.@preserve colentries check(ccall((:grmp_space_create, lib), Cint,
            (Ptr{Cvoid}, Cint, Cint, Int64, Cint, Ptr{Int32}, Ref{Ptr{Cvoid}}),
            g.h, fecode(FEType), get_ncomponents(FEType), FES.ndofs, size(colentries, 1), colentries, h))
        d = DSpace(h[], g)
        finalizer(x -> destroy(:grmp_space_destroy, x), d)
        d
    end
end

const BSPACES = WeakKeyDict{Any,DSpace}()
function device_space(FES::FESpace{Float64,Int32,FEType}, ::Type{ON_BFACES}) where {FEType}
    get!(BSPACES, FES) do
        g = device_grid(FES.xgrid, ON_BFACES)
        nb = num_sources(FES.xgrid[BFaceNodes])
        colentries = Matrix{Int32}(reshape(FES[BFaceDofs].colentries, :, nb))       # Dofmap4AssemblyType(ON_BFACES), dofmaps.jl:45
        h = Ref{Ptr{Cvoid}}()
        GC.@preserve colentries check(ccall((:grmp_space_create, lib), Cint,
            (Ptr{Cvoid}, Cint, Cint, Int64, Cint, Ptr{Int32}, Ref{Ptr{Cvoid}}),
            g.h, fecode(FEType), get_ncomponents(FEType), FES.ndofs, size(colentries, 1), colentries, h))
        d = DSpace(h[], g)
        finalizer(x -> destroy(:grmp_space_destroy, x), d)
        d
    end
end
device_space(FES::FESpace{Float64,Int32}, ::Type{ON_CELLS}) = device_space(FES)
const DeviceAT = Union{ON_CELLS,ON_BFACES}
# ON_BFACES: Identity of H1P1 / H1P2 and NormalFlux of HDIVRT0 / HDIVBDM1 (the best-approximation boundary data of
# boundarydata.jl:297-347); everything else on boundary faces (TangentFlux, face bubbles of H1BR, NormalFlux of H1 spaces) stays
# with the reference
bface_ok(AP) = (all(F -> eltype(F) <: Union{H1P1,H1P2} && !F.broken, AP.FES) && all(o -> o == Identity, AP.operators)) ||
               (all(F -> eltype(F) <: Union{HDIVRT0,HDIVBDM1} && !F.broken, AP.FES) && all(o -> o == NormalFlux, AP.operators))

# tables straight out of the reference's FEEvaluator (feevaluator.jl:34-138): ForwardDiff's bits travel unchanged
function evaltab(ev)   # ev::GRMP.SingleFEEvaluator
    vals = permutedims(cat(ev.refbasisvals...; dims = 3), (2, 1, 3))[:]     # refbasisvals[i][dof, comp] -> memory [i][dof][comp]
    ders = ev.derivorder > 0 ? ev.refbasisderivvals[:] : Float64[]         # [dof + comp nd_all, j, i], column-major == [i][j][row]
    return vals, ders, EvalTab(size(ev.refbasisvals[1], 1), size(ev.refbasisvals[1], 2), pointer(vals), isempty(ders) ? C_NULL : pointer(ders))
end

# ---- what is on the device path -----------------------------------------------------------------------------------------
"""
    hooke_parameters(action) -> (actcode, [μ, λ]) or nothing

The library evaluates `NoAction` and the constant isotropic Hooke tensors of HookStiffnessOperator2D/3D
(pdeoperators.jl:256-273, 296-315).  Their kernels are closures over (μ, λ); instead of reaching into the closure the
tensor is PROBED: C e_i for the unit vectors, and accepted only if it has exactly the Hooke structure
(c11 = λ + 2μ, c12 = λ, c33/c44 = μ, zeros elsewhere) -- any other user Action returns `nothing` and stays on the reference path.
"""
function hooke_parameters(action)
    action isa NoAction && return (0, Float64[])
    action isa GRMP.DefaultUserAction || return nothing
    (GRMP.is_xdependent(action) || GRMP.is_timedependent(action) || GRMP.is_itemdependent(action) || GRMP.is_xrefdependent(action)) && return nothing
    n = action.argsizes[1]
    (n == action.argsizes[2] && (n == 3 || n == 6)) || return nothing
    C = zeros(n, n); r = zeros(n)
    for i = 1:n
        e = zeros(n); e[i] = 1
        fill!(r, 0); action.kernel(r, e); C[:, i] = r
    end
    nd = n == 3 ? 2 : 3
    μ, λ = C[n, n], C[1, 2]
    for i = 1:n, j = 1:n
        expect = i == j ? (i <= nd ? λ + 2μ : μ) : (i <= nd && j <= nd ? λ : 0.0)
        C[i, j] == expect || return nothing
    end
    # linearity spot check (a kernel that is not a constant tensor must not pass)
    x = collect(1.0:n); action.kernel(r, x)
    maximum(abs.(r .- C * x)) <= 1e-12 * maximum(abs.(C)) * n || return nothing
    return (n == 3 ? 1 : 2, Float64[μ, λ])
end

struct BlfPlan; act::Int; par::Vector{Float64}; ops::Tuple{Int,Int}; end
function blf_plan(AP::AssemblyPattern{APT,Tv,AT}) where {APT,Tv,AT}
    length(AP.FES) == 2 && length(AP.operators) == 2 || return nothing
    AT <: ON_BFACES && !(bface_ok(AP) && AP.action isa NoAction) && return nothing
    AP.FES[1].xgrid === AP.FES[2].xgrid || return nothing
    AP.FES[1].xgrid[UniqueCellGeometries] in ([Triangle2D], [Tetrahedron3D]) || return nothing
    all(F -> fecode(eltype(F)) !== nothing && !F.broken || eltype(F) <: L2P0, AP.FES) || return nothing
    o1, o2 = opcode(AP.operators[1]), opcode(AP.operators[2])
    (o1 === nothing || o2 === nothing) && return nothing
    hp = hooke_parameters(AP.action)
    hp === nothing && return nothing
    hp[1] != 0 && AP.apply_action_to != [1] && return nothing
    return BlfPlan(hp[1], hp[2], (o1, o2))
end

# ---- BilinearForm -----------------------------------------------------------------------------------------------------
# keyed by the pattern AND the output orientation; the factor of the symbolic pass is remembered: the reference's pattern is
# value dependent (_addnz skips exact zeros, fematrix.jl:54-65), so a frozen pattern is reused only with skip_preps = true
# (the reference's own contract for reassembly, solvers.jl:556) or when nothing that defines it changed.
const PATTERNS = WeakKeyDict{Any,Dict{Bool,DBlf}}()

function build_blf(A, AP::AssemblyPattern{APT,Tv,AT}, plan::BlfPlan, factor, transposed_assembly) where {APT,Tv,AT}
    e1 = GRMP.get_basisevaler(AP.AM, 1, 1); e2 = GRMP.get_basisevaler(AP.AM, 2, 1)
    v1, d1, t1 = evaltab(e1); v2, d2, t2 = evaltab(e2)
    w = Vector{Float64}(GRMP.get_qweights(AP.AM))
    par = plan.par
    regions = AP.regions == [0] ? Int32[] : Vector{Int32}(AP.regions)
    s1, s2 = device_space(AP.FES[1], AT), device_space(AP.FES[2], AT)
    h = Ref{Ptr{Cvoid}}()
    GC.@preserve v1 d1 v2 d2 w par regions check(ccall((:grmp_blf_create, lib), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float64}, Cint, Cint, Ptr{Int32}, Cint, Cint, Ptr{Float64}, Ref{EvalTab}, Ref{EvalTab}, Ref{Ptr{Cvoid}}),
        s1.h, s2.h, plan.ops[1], plan.ops[2], plan.act, isempty(par) ? C_NULL : pointer(par), aptcode(APT), transposed_assembly,
        isempty(regions) ? C_NULL : pointer(regions), length(regions), length(w), w, t1, t2, h))
    b = DBlf(h[], 0, Int64[], Int64[], Float64(factor), transposed_assembly, (s1, s2))
    finalizer(x -> destroy(:grmp_blf_destroy, x), b)
    symbolic!(b, A, factor)
    return b
end
function symbolic!(b::DBlf, A, factor)
    nnz = Ref{Int64}(0)
    check(ccall((:grmp_blf_symbolic, lib), Cint, (Ptr{Cvoid}, Float64, Ref{Int64}), b.h, factor, nnz))
    b.nnz = nnz[]; b.factor = factor
    b.colptr = Vector{Int64}(undef, size(A, 2) + 1); b.rowval = Vector{Int64}(undef, nnz[])
    check(ccall((:grmp_blf_get_pattern, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), b.h, b.colptr, b.rowval))
end

"""
    assemble!(A::FEMatrixBlock, AP; factor, factor_transpose, skip_preps, transposed_assembly, transpose_copy)   (device version)

skip_preps = false (first assembly, bilinearform.jl:104-107): prepare_assembly! on the host (unchanged), then the symbolic
pass runs (again) with this `factor` -- grmp_blf_create once per (pattern, orientation), grmp_blf_symbolic + pattern download.
skip_preps = true: numeric assembly on the frozen pattern (grmp_blf_numeric).  The block is then installed into `A.entries`
(single-block matrix with an empty target: `A.entries.cscmatrix = SparseMatrixCSC(m, n, colptr, rowval, nzval)`; otherwise
an addblock!-style merge), the transposed copy likewise.
"""
function GRMP.assemble!(A::FEMatrixBlock{Float64,Int64,Float64,Int32}, AP::AssemblyPattern{APT,Float64,AT,Float64,Int32};
        factor = 1, factor_transpose = factor, skip_preps::Bool = false, fixed_arguments = nothing,
        transposed_assembly::Bool = false, transpose_copy = nothing) where {APT<:GRMP.APT_BilinearForm,AT<:DeviceAT}
    plan = blf_plan(AP)
    if plan === nothing || !(transpose_copy === nothing || transpose_copy isa FEMatrixBlock{Float64,Int64,Float64,Int32})
        # not on the device path: the reference's own loop, all keywords forwarded
        return invoke(GRMP.assemble!, Tuple{AbstractArray{Float64,2},AssemblyPattern{APT,Float64,AT,Float64,Int32}}, A, AP;
                      factor, factor_transpose, skip_preps, fixed_arguments, transposed_assembly, transpose_copy)
    end
    skip_preps || GRMP.prepare_assembly!(AP)
    tr = transposed_assembly && !(APT <: GRMP.APT_SymmetricBilinearForm)
    byor = get!(() -> Dict{Bool,DBlf}(), PATTERNS, AP)
    d = get(byor, tr, nothing)
    if d === nothing
        d = byor[tr] = build_blf(A, AP, plan, Float64(factor), tr)
    elseif !skip_preps
        symbolic!(d, A, Float64(factor))           # prepare_assembly! ran again: quadrature / tables / factor may define a new pattern
    end
    nzval = Vector{Float64}(undef, d.nnz)
    check(ccall((:grmp_blf_numeric, lib), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}), d.h, Float64(factor), nzval))
    install_block!(A, SparseMatrixCSC(size(A, 1), size(A, 2), d.colptr, d.rowval, nzval))
    if transpose_copy !== nothing
        m = size(transpose_copy, 2)       # = rows of A
        cpt = Vector{Int64}(undef, m + 1); rvt = Vector{Int64}(undef, d.nnz); nzt = Vector{Float64}(undef, d.nnz)
        check(ccall((:grmp_blf_transpose_copy, lib), Cint, (Ptr{Cvoid}, Float64, Float64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
            d.h, Float64(factor), Float64(factor_transpose), cpt, rvt, nzt))
        install_block!(transpose_copy, SparseMatrixCSC(size(transpose_copy, 1), m, cpt, rvt, nzt))
    end
    AP.last_allocations = 0
    return nothing
end

"""
    assemble_from_host!(A, AP; factor)

Reassembly after the grid moved (same topology): one `grmp_blf_assemble_host` call uploads `Coordinates` and `CellVolumes`,
assembles on the frozen pattern and downloads `nzval`.  `AP` must have been assembled once through `assemble!` above.
"""
function assemble_from_host!(A::FEMatrixBlock{Float64,Int64,Float64,Int32}, AP::AssemblyPattern; factor = 1, transposed_assembly::Bool = false)
    d = PATTERNS[AP][transposed_assembly]
    xgrid = AP.FES[1].xgrid
    coords = xgrid[Coordinates]; vol = Vector{Float64}(xgrid[CellVolumes])
    nzval = Vector{Float64}(undef, d.nnz)
    GC.@preserve coords vol check(ccall((:grmp_blf_assemble_host, lib), Cint,
        (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
        d.h, Float64(factor), coords, vol, C_NULL, C_NULL, C_NULL, nzval))
    install_block!(A, SparseMatrixCSC(size(A, 1), size(A, 2), d.colptr, d.rowval, nzval))
    return nothing
end

# single-block FEMatrix with an empty target: adopt the CSC; otherwise merge entry by entry like _addnz (explicit zeros of
# the block are kept: they are entries of the reference's pattern, fematrix.jl:54-65 only skips zero CONTRIBUTIONS)
function install_block!(A::FEMatrixBlock, B::SparseMatrixCSC{Float64,Int64})
    E = A.entries
    flush!(E)
    if A.offsetX == 0 && A.offsetY == 0 && size(E) == size(B) && nnz(E.cscmatrix) == 0
        E.cscmatrix = B
    else
        rows = rowvals(B); vals = nonzeros(B)
        for j = 1:size(B, 2), k in nzrange(B, j)
            rawupdateindex!(E, +, vals[k], rows[k] + A.offsetX, j + A.offsetY)
        end
        flush!(E)
    end
end

# ---- device-resident follow-ups (SURVEY.md 8f N1 / N3): the matrix of the last assemble! stays on the GPU ---------------------
"`matmul!(a, AP, b; factor, transposed)`: a += A b factor on the device matrix of `AP` (addblock_matmul!, fematrix.jl:402-473)"
function matmul!(a::Vector{Float64}, AP::AssemblyPattern, b::Vector{Float64}; factor = 1, transposed::Bool = false, transposed_assembly::Bool = false)
    d = PATTERNS[AP][transposed_assembly]
    check(ccall((:grmp_blf_matmul, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Cint), d.h, b, a, Float64(factor), transposed))
    return a
end
"`residual!(r, AP, x, b, fixed_dofs)`: r = A x - b, r[fixed_dofs] = 0, returns sum r_i^2 (solve_direct!'s check, solvers.jl:661-668)"
function residual!(r::Vector{Float64}, AP::AssemblyPattern, x::Vector{Float64}, b::Vector{Float64}, fixed_dofs::Vector{Int64}; transposed_assembly::Bool = false)
    d = PATTERNS[AP][transposed_assembly]
    n2 = Ref{Float64}(0)
    check(ccall((:grmp_blf_residual, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Int64, Ptr{Float64}, Ref{Float64}),
        d.h, x, b, fixed_dofs, length(fixed_dofs), r, n2))
    return n2[]
end
"`apply_penalties!(AP, fixed_dofs, penalty)` on the device values (fematrix.jl:349-355); throws if a fixed dof has no stored diagonal"
function apply_penalties!(AP::AssemblyPattern, fixed_dofs::Vector{Int64}, penalty; transposed_assembly::Bool = false)
    d = PATTERNS[AP][transposed_assembly]
    missing_diag = Ref{Int64}(0)
    check(ccall((:grmp_blf_apply_penalties, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Float64, Ref{Int64}), d.h, fixed_dofs, length(fixed_dofs), Float64(penalty), missing_diag))
end

# ---- LinearForm ------------------------------------------------------------------------------------------------------
struct LfPlan; op::Int; fsrc::Int; end
function lf_plan(AP::AssemblyPattern{APT,Tv,AT}) where {APT,Tv,AT}
    length(AP.FES) == 1 && length(AP.operators) == 1 || return nothing
    AT <: ON_BFACES && !bface_ok(AP) && return nothing
    F = AP.FES[1]
    F.xgrid[UniqueCellGeometries] in ([Triangle2D], [Tetrahedron3D]) || return nothing
    (fecode(eltype(F)) !== nothing && (!F.broken || eltype(F) <: L2P0)) || return nothing
    o = opcode(AP.operators[1])
    o === nothing && return nothing
    a = AP.action
    a isa NoAction && return LfPlan(o, F_NONE)
    a isa GRMP.DefaultUserAction || return nothing
    a.argsizes[2] == 0 || return nothing                                  # fdot_action(f) / fdotn_action(f): the result does not depend on an input (actions.jl:119-192)
    GRMP.is_xrefdependent(a) && return nothing                            # the reference's LinearForm loop never sets action.xref: "L" kernels stay there
    # x- and item-dependent kernels are tabulated on the host item by item (tabulate_action sets action.x / action.item the way the
    # reference loop does, linearform.jl:172-201); only a kernel without any dependency is a constant
    return LfPlan(o, (GRMP.is_xdependent(a) || GRMP.is_itemdependent(a)) ? F_QP_TABLE : F_CONST)
end

const LFS = WeakKeyDict{Any,DLf}()

# f at the quadrature points of every cell, evaluated exactly as the reference loop does it (linearform.jl:197-201:
# update_trafo! / eval_trafo!(action.x, L2G, xref[i]); eval_action!(action, input)) -> table [ncells][nq][resultdim]
function tabulate_action(AP::AssemblyPattern{APT,Tv,AT}, ev, nq::Int) where {APT,Tv,AT}
    action = AP.action
    rd = action.argsizes[1]
    xgrid = AP.FES[1].xgrid
    ncells = num_sources(xgrid[GRMP.GridComponentNodes4AssemblyType(AT)])       # items: cells, or boundary faces (ev.L2G is the transformer of AT)
    regions = AP.regions; allitems = regions == [0]
    xreg = xgrid[GRMP.GridComponentRegions4AssemblyType(AT)]
    input = zeros(Float64, 0)
    xdep, idep = GRMP.is_xdependent(action), GRMP.is_itemdependent(action)
    if !(xdep || idep)
        GRMP.eval_action!(action, input)
        return Vector{Float64}(action.val[1:rd])
    end
    tab = zeros(Float64, rd, nq, ncells)
    for cell = 1:ncells
        (allitems || xreg[cell] in regions) || continue
        if idep                                   # linearform.jl:172-177 (continuous operators: the dofitem is the item, di = 1)
            action.item[1] = cell; action.item[2] = cell; action.item[3] = xreg[cell]; action.item[4] = 1
        end
        xdep && update_trafo!(ev.L2G, cell)
        for i = 1:nq
            xdep && eval_trafo!(action.x, ev.L2G, ev.xref[i])
            GRMP.eval_action!(action, input)
            @views tab[:, i, cell] .= action.val[1:rd]
        end
    end
    return vec(tab)
end

"""
    assemble!(b::FEVectorBlock, AP; skip_preps, factor)   (device version; reference: linearform.jl:47-237, 239-251)

b[dof + b.offset] += contributions in cell order.  A DataFunction cannot cross the C ABI: it is tabulated on the host at the
quadrature points (GRMP_F_QP_TABLE) or passed as a constant (GRMP_F_CONST); the basis evaluation, contraction and scatter
run on the device.
"""
function GRMP.assemble!(b::FEVectorBlock{Float64,Float64,Int32}, AP::AssemblyPattern{APT,Float64,AT,Float64,Int32};
        skip_preps::Bool = false, factor = 1, fixed_arguments = nothing) where {APT<:GRMP.APT_LinearForm,AT<:DeviceAT}
    plan = lf_plan(AP)
    if plan === nothing
        return invoke(GRMP.assemble!, Tuple{Union{AbstractArray{Float64,1},AbstractArray{Float64,2}},AssemblyPattern{APT,Float64,AT,Float64,Int32}},
                      b, AP; skip_preps, factor, fixed_arguments)
    end
    @assert b.FES == AP.FES[1]
    return lf_device!(b.entries, AP, plan, skip_preps, Float64(factor), b.offset)
end
# the plain-vector call of boundarydata.jl:315 (`assemble!(b, RHS_bnd)` with b::Array{T,1}): boundary forms only, so that the
# method table of cell forms on plain vectors stays the reference's
function GRMP.assemble!(b::Vector{Float64}, AP::AssemblyPattern{APT,Float64,ON_BFACES,Float64,Int32};
        skip_preps::Bool = false, factor = 1, fixed_arguments = nothing, offset = 0) where {APT<:GRMP.APT_LinearForm}
    plan = lf_plan(AP)
    if plan === nothing
        return invoke(GRMP.assemble!, Tuple{Union{AbstractArray{Float64,1},AbstractArray{Float64,2}},AssemblyPattern{APT,Float64,ON_BFACES,Float64,Int32}},
                      b, AP; skip_preps, factor, fixed_arguments, offset)
    end
    return lf_device!(b, AP, plan, skip_preps, Float64(factor), offset)
end

function lf_device!(entries::Vector{Float64}, AP::AssemblyPattern{APT,Float64,AT}, plan::LfPlan, skip_preps::Bool, factor::Float64, offset) where {APT,AT}
    skip_preps || GRMP.prepare_assembly!(AP)
    ev = GRMP.get_basisevaler(AP.AM, 1, 1)
    w = Vector{Float64}(GRMP.get_qweights(AP.AM))
    if !skip_preps && haskey(LFS, AP)      # prepare_assembly! ran again: the tables may have changed
        destroy(:grmp_lf_destroy, LFS[AP]); delete!(LFS, AP)
    end
    d = get!(LFS, AP) do
        v, dv, t = evaltab(ev)
        regions = AP.regions == [0] ? Int32[] : Vector{Int32}(AP.regions)
        s = device_space(AP.FES[1], AT)
        h = Ref{Ptr{Cvoid}}()
        GC.@preserve v dv w regions check(ccall((:grmp_lf_create, lib), Cint,
            (Ptr{Cvoid}, Cint, Ptr{Int32}, Cint, Cint, Ptr{Float64}, Ref{EvalTab}, Ref{Ptr{Cvoid}}),
            s.h, plan.op, isempty(regions) ? C_NULL : pointer(regions), length(regions), length(w), w, t, h))
        l = DLf(h[], s)
        finalizer(x -> destroy(:grmp_lf_destroy, x), l)
        l
    end
    fdata = plan.fsrc == F_NONE ? Float64[] : tabulate_action(AP, ev, length(w))
    GC.@preserve fdata entries check(ccall((:grmp_lf_assemble, lib), Cint,
        (Ptr{Cvoid}, Float64, Cint, Ptr{Float64}, Ptr{Float64}, Int64),
        d.h, factor, plan.fsrc, isempty(fdata) ? C_NULL : pointer(fdata), entries, offset))
    AP.last_allocations = 0
    return nothing
end

# ---- trilinear forms with one coefficient argument (SURVEY.md 8f N4, first slice) ---------------------------------------------------
# assemble!(A, AP, FEB; fixed_arguments = [1]) with three FESpaces (bilinearform.jl:235-257), as `assemble_operator!` calls it for a
# ConvectionOperator(a_from, a_operator, xdim, ncomponents; a_to = 1) (pdeoperators.jl:435-510, 986-987).  The kernel is an opaque closure:
# it is PROBED on random inputs and accepted only if it is result[j] = sum_k input[k] * input[xdim + (j-1) xdim + k].
function is_convection_kernel(action, xdim::Int, nc::Int)
    action isa GRMP.DefaultUserAction || return false
    (GRMP.is_xdependent(action) || GRMP.is_timedependent(action) || GRMP.is_itemdependent(action) || GRMP.is_xrefdependent(action)) && return false
    action.argsizes[1] == nc && action.argsizes[2] == xdim + nc * xdim || return false
    r = zeros(nc)
    for trial = 1:3
        x = [sin(1.0 + 0.7 * i * trial) for i = 1:xdim+nc*xdim]
        fill!(r, 0); action.kernel(r, x)
        for j = 1:nc
            e = 0.0
            for k = 1:xdim; e += x[k] * x[xdim+(j-1)*xdim+k]; end
            r[j] == e || return false
        end
    end
    return true
end

const TRIPATTERNS = WeakKeyDict{Any,Dict{Bool,DBlf}}()

function GRMP.assemble!(A::FEMatrixBlock{Float64,Int64,Float64,Int32}, AP::AssemblyPattern{APT,Float64,ON_CELLS,Float64,Int32},
        FEB::Array{<:FEVectorBlock{Float64,Float64,Int32},1};
        factor = 1, factor_transpose = factor, skip_preps::Bool = false, fixed_arguments = nothing,
        transposed_assembly::Bool = false, transpose_copy = nothing) where {APT<:GRMP.APT_BilinearForm}
    fallback() = invoke(GRMP.assemble!, Tuple{FEMatrixBlock,AssemblyPattern{APT,Float64,ON_CELLS},Array{<:FEVectorBlock{Float64,Float64,Int32},1}},
                        A, AP, FEB; factor, factor_transpose, skip_preps, fixed_arguments, transposed_assembly, transpose_copy)
    (length(AP.FES) == 3 && length(FEB) == 1 && transpose_copy === nothing && APT === GRMP.APT_BilinearForm &&
     (fixed_arguments === nothing || fixed_arguments == [1])) || return fallback()
    xdim = size(AP.FES[1].xgrid[Coordinates], 1)
    oa, o1, o2 = opcode(AP.operators[1]), opcode(AP.operators[2]), opcode(AP.operators[3])
    (oa === nothing || o1 === nothing || o2 === nothing || any(F -> fecode(eltype(F)) === nothing, AP.FES)) && return fallback()
    nc = GRMP.Length4Operator(AP.operators[3], xdim, get_ncomponents(eltype(AP.FES[3])))
    is_convection_kernel(AP.action, GRMP.Length4Operator(AP.operators[1], xdim, get_ncomponents(eltype(AP.FES[1]))), nc) || return fallback()
    skip_preps || GRMP.prepare_assembly!(AP)
    byor = get!(() -> Dict{Bool,DBlf}(), TRIPATTERNS, AP)
    d = get(byor, transposed_assembly, nothing)
    fresh = d === nothing
    if fresh
        e1 = GRMP.get_basisevaler(AP.AM, 2, 1); e2 = GRMP.get_basisevaler(AP.AM, 3, 1)
        v1, d1, t1 = evaltab(e1); v2, d2, t2 = evaltab(e2)
        w = Vector{Float64}(GRMP.get_qweights(AP.AM))
        regions = AP.regions == [0] ? Int32[] : Vector{Int32}(AP.regions)
        s1, s2 = device_space(AP.FES[2]), device_space(AP.FES[3])
        h = Ref{Ptr{Cvoid}}()
        GC.@preserve v1 d1 v2 d2 w regions check(ccall((:grmp_blf_create, lib), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float64}, Cint, Cint, Ptr{Int32}, Cint, Cint, Ptr{Float64}, Ref{EvalTab}, Ref{EvalTab}, Ref{Ptr{Cvoid}}),
            s1.h, s2.h, o1, o2, 3, C_NULL, 0, transposed_assembly, isempty(regions) ? C_NULL : pointer(regions), length(regions), length(w), w, t1, t2, h))
        d = byor[transposed_assembly] = DBlf(h[], 0, Int64[], Int64[], Float64(factor), transposed_assembly, (s1, s2))
        finalizer(x -> destroy(:grmp_blf_destroy, x), d)
    end
    # the coefficient function: FEB[1] evaluated at the quadrature points on the device
    ea = GRMP.get_basisevaler(AP.AM, 1, 1)
    va, da, ta = evaltab(ea)
    coeffs = FEB[1].entries[FEB[1].offset+1:FEB[1].last_index]
    keep = skip_preps && !fresh
    GC.@preserve va da coeffs check(ccall((:grmp_blf_set_fixed_argument, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ref{EvalTab}, Ptr{Float64}, Cint),
        d.h, device_space(AP.FES[1]).h, oa, ta, coeffs, keep))
    keep || symbolic!(d, A, Float64(factor))
    nzval = Vector{Float64}(undef, d.nnz)
    check(ccall((:grmp_blf_numeric, lib), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}), d.h, Float64(factor), nzval))
    install_block!(A, SparseMatrixCSC(size(A, 1), size(A, 2), d.colptr, d.rowval, nzval))
    AP.last_allocations = 0
    return nothing
end

# ---- NonlinearForm: Newton form of the convection term (SURVEY.md 8f N4) -------------------------------------------------------------
# full_assemble!(A, b, AP, FEB; factor, transposed_assembly, skip_preps) (nonlinearform.jl:44-245) for the pattern that
# ConvectionOperator(a_from, a_operator, xdim, ncomponents; newton = true) creates (pdeoperators.jl:459-493): the kernel AND the user
# jacobian are probed (value[j] = sum_k in[k] in[xdim+(j-1)xdim+k]; jac[j,k] = in[xdim+(j-1)xdim+k], jac[j,xdim+(j-1)xdim+k] = in[k]).
function is_newton_convection(h, xdim::Int, nc::Int)
    h isa GRMP.OperatorWithUserJacobian || return false
    n = xdim + nc * xdim
    (h.argsizes[1] == nc && h.argsizes[2] == n) || return false
    x = [cos(0.3 + 0.9 * i) for i = 1:n]
    GRMP.eval_jacobian!(h, x)
    J = Matrix(h.jac)
    for j = 1:nc
        e = 0.0
        for k = 1:xdim; e += x[k] * x[xdim+(j-1)*xdim+k]; end
        h.val[j] == e || return false
        for c = 1:n
            expect = c <= xdim ? x[xdim+(j-1)*xdim+c] : (xdim + (j - 1) * xdim < c <= xdim + j * xdim ? x[c-xdim-(j-1)*xdim] : 0.0)
            J[j, c] == expect || return false
        end
    end
    return true
end

const NLPATTERNS = WeakKeyDict{Any,DBlf}()

function GRMP.full_assemble!(A::FEMatrixBlock{Float64,Int64,Float64,Int32}, b::FEVectorBlock{Float64,Float64,Int32},
        AP::AssemblyPattern{APT,Float64,ON_CELLS,Float64,Int32}, FEB::Array{<:FEVectorBlock{Float64,Float64,Int32},1};
        factor = 1, transposed_assembly::Bool = false, skip_preps::Bool = false) where {APT<:GRMP.APT_NonlinearForm}
    fallback() = invoke(GRMP.full_assemble!, Tuple{FEMatrixBlock,FEVectorBlock,AssemblyPattern{APT,Float64,ON_CELLS},Array{<:FEVectorBlock{Float64,Float64,Int32},1}},
                        A, b, AP, FEB; factor, transposed_assembly, skip_preps)
    (length(AP.FES) == 3 && AP.FES[1] === AP.FES[2] === AP.FES[3] && AP.newton_args == [1, 2] && FEB[1] === FEB[2]) || return fallback()
    xdim = size(AP.FES[1].xgrid[Coordinates], 1)
    oa, og, ot = opcode(AP.operators[1]), opcode(AP.operators[2]), opcode(AP.operators[3])
    (oa === nothing || og === nothing || ot === nothing || fecode(eltype(AP.FES[1])) === nothing) && return fallback()
    nc = GRMP.Length4Operator(AP.operators[3], xdim, get_ncomponents(eltype(AP.FES[3])))
    is_newton_convection(AP.action, GRMP.Length4Operator(AP.operators[1], xdim, get_ncomponents(eltype(AP.FES[1]))), nc) || return fallback()
    skip_preps || GRMP.prepare_assembly!(AP)
    fresh = !haskey(NLPATTERNS, AP)
    d = get!(NLPATTERNS, AP) do
        e1 = GRMP.get_basisevaler(AP.AM, 2, 1); e2 = GRMP.get_basisevaler(AP.AM, 3, 1)
        v1, d1, t1 = evaltab(e1); v2, d2, t2 = evaltab(e2)
        w = Vector{Float64}(GRMP.get_qweights(AP.AM))
        regions = AP.regions == [0] ? Int32[] : Vector{Int32}(AP.regions)
        su = device_space(AP.FES[1])
        h = Ref{Ptr{Cvoid}}()
        GC.@preserve v1 d1 v2 d2 w regions check(ccall((:grmp_blf_create, lib), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float64}, Cint, Cint, Ptr{Int32}, Cint, Cint, Ptr{Float64}, Ref{EvalTab}, Ref{EvalTab}, Ref{Ptr{Cvoid}}),
            su.h, su.h, og, ot, 4, C_NULL, 0, transposed_assembly, isempty(regions) ? C_NULL : pointer(regions), length(regions), length(w), w, t1, t2, h))
        x = DBlf(h[], 0, Int64[], Int64[], Float64(factor), transposed_assembly, (su, su))
        finalizer(y -> destroy(:grmp_blf_destroy, y), x)
        x
    end
    ea = GRMP.get_basisevaler(AP.AM, 1, 1)
    va, da, ta = evaltab(ea)
    coeffs = FEB[1].entries[FEB[1].offset+1:FEB[1].last_index]
    keep = skip_preps && !fresh
    GC.@preserve va da coeffs check(ccall((:grmp_blf_set_newton_argument, lib), Cint, (Ptr{Cvoid}, Cint, Ref{EvalTab}, Ptr{Float64}, Cint), d.h, oa, ta, coeffs, keep))
    keep || symbolic!(d, A, Float64(factor))
    nzval = Vector{Float64}(undef, d.nnz)
    check(ccall((:grmp_blf_numeric, lib), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}), d.h, Float64(factor), nzval))
    install_block!(A, SparseMatrixCSC(size(A, 1), size(A, 2), d.colptr, d.rowval, nzval))
    entries = b.entries
    GC.@preserve entries check(ccall((:grmp_blf_newton_rhs, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), d.h, entries, b.offset))
    AP.last_allocations = 0
    return nothing
end

# ---- ItemIntegrator (src/assemblypatterns/itemintegrator.jl:160-360), one argument -------------------------------------------------
# The library evaluates NoAction and the kernels of L2NormIntegrator / L2ErrorIntegrator; an ItemIntegrator carries them as opaque
# closures, so the device versions are constructed explicitly:
#     II = GRMPCuda.L2ErrorIntegrator(u_exact, Identity; quadorder = 4)       # same arguments as the reference constructor
#     err2 = GRMPCuda.evaluate(II, Solution[1])                               # = evaluate(L2ErrorIntegrator(...), Solution[1])
struct DeviceItemIntegrator
    operator::DataType
    kind::Int                      # GRMP_II_NONE / L2NORM / L2ERROR
    data::Union{Nothing,GRMP.AbstractUserDataType}
    factor::Float64
    bonus_quadorder::Int
    regions::Vector{Int}
    AT::DataType                   # ON_CELLS, or ON_BFACES (boundary integrals: Identity of H1P1 / H1P2, NormalFlux of HDIVRT0 / HDIVBDM1)
end
ItemIntegrator(operator::DataType; AT = ON_CELLS, regions = [0]) = DeviceItemIntegrator(operator, 0, nothing, 1.0, 0, regions, AT)
L2NormIntegrator(ncomponents::Int, operator::DataType; AT = ON_CELLS, quadorder = 2, regions = [0]) =
    DeviceItemIntegrator(operator, 1, nothing, 1.0, quadorder, regions, AT)
L2ErrorIntegrator(compare_data, operator::DataType = Identity; AT = ON_CELLS, quadorder = "auto", factor = 1, regions = [0]) =
    DeviceItemIntegrator(operator, 2, compare_data, Float64(factor), quadorder == "auto" ? 2 * compare_data.bonus_quadorder : quadorder, regions, AT)

function prepare(II::DeviceItemIntegrator, FES::FESpace{Float64,Int32,FEType}) where {FEType}
    xgrid = FES.xgrid
    EG = xgrid[GRMP.GridComponentUniqueGeometries4AssemblyType(II.AT)][1]           # cells, or the boundary-face geometry
    order = max(II.bonus_quadorder + GRMP.get_polynomialorder(FEType, EG) + GRMP.QuadratureOrderShift4Operator(II.operator), 0)   # assemblypatterns.jl:559-565
    qf = QuadratureRule{Float64,EG}(order)
    ev = FEEvaluator(FES, II.operator, qf; AT = II.AT)
    v, dv, t = evaltab(ev)
    w = Vector{Float64}(qf.w)
    regions = II.regions == [0] ? Int32[] : Vector{Int32}(II.regions)
    h = Ref{Ptr{Cvoid}}()
    GC.@preserve v dv w regions check(ccall((:grmp_ii_create, lib), Cint,
        (Ptr{Cvoid}, Cint, Cint, Ptr{Int32}, Cint, Cint, Ptr{Float64}, Ref{EvalTab}, Ref{Ptr{Cvoid}}),
        device_space(FES, II.AT).h, opcode(II.operator), II.kind, isempty(regions) ? C_NULL : pointer(regions), length(regions), length(w), w, t, h))
    return h[], ev, qf
end

# compare_data at the quadrature points, evaluated like L2error_function does (itemintegrator.jl:52-69) -> [resultdim, nq, ncells]
function tabulate_data(data, ev, qf, xgrid, AT = ON_CELLS)
    ncells = num_sources(xgrid[GRMP.GridComponentNodes4AssemblyType(AT)]); rd = data.argsizes[1]
    tab = zeros(Float64, rd, length(qf.w), ncells)
    x = zeros(Float64, size(xgrid[Coordinates], 1))
    for cell = 1:ncells
        update_trafo!(ev.L2G, cell)
        for i = 1:length(qf.w)
            eval_trafo!(x, ev.L2G, ev.xref[i])
            if GRMP.is_xdependent(data); data.x = x; end
            GRMP.eval_data!(data)
            @views tab[:, i, cell] .= data.val[1:rd]
        end
    end
    return vec(tab)
end

"`evaluate!(b, II, FEB)`: b[j, item] += ... (itemintegrator.jl:160-300); `evaluate(II, FEB)`: the accumulation over all items (316-360)"
function evaluate!(b::Matrix{Float64}, II::DeviceItemIntegrator, FEB::FEVectorBlock{Float64,Float64,Int32}; total = nothing)
    h, ev, qf = prepare(II, FEB.FES)
    try
        data = II.kind == 2 ? tabulate_data(II.data, ev, qf, FEB.FES.xgrid, II.AT) : Float64[]
        coeffs = FEB.entries[FEB.offset+1:FEB.last_index]
        GC.@preserve data coeffs b check(ccall((:grmp_ii_evaluate, lib), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            h, coeffs, II.factor, isempty(data) ? C_NULL : pointer(data), isempty(b) ? C_NULL : pointer(b), total === nothing ? C_NULL : pointer(total)))
    finally
        ccall((:grmp_ii_destroy, lib), Cint, (Ptr{Cvoid},), h)
    end
    return nothing
end
function evaluate(II::DeviceItemIntegrator, FEB::FEVectorBlock{Float64,Float64,Int32})
    rd = Ref{Cint}(0)
    total = zeros(Float64, 16)
    evaluate!(Matrix{Float64}(undef, 0, 0), II, FEB; total = total)
    n = II.kind == 0 ? GRMP.Length4Operator(II.operator, size(FEB.FES.xgrid[Coordinates], 1), get_ncomponents(eltype(FEB.FES))) : 1
    return n == 1 ? total[1] : total[1:n]
end

end # module
