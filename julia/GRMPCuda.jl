# GRMPCuda.jl -- Julia glue that puts libgrmp_cuda behind the unchanged
# GradientRobustMultiPhysics.jl API (PDEDescription / add_operator! / assemble! / solve!).
#
# NOT RUNNABLE IN THE BUILD CONTAINER (no julia binary, no network).  It is kept tiny and
# mechanical: every `ccall` below has a line-for-line ctypes twin in
# gradientrobustmultiphysics.jl_b200/_lib.py + assembly.py, which is what the test-suite runs.
#
# How it hooks in: it adds *more specific* methods of
#     assemble!(A::FEMatrixBlock, AP::AssemblyPattern{<:APT_BilinearForm,Float64,ON_CELLS}, FEB; ...)
#     assemble!(b::FEVectorBlock, AP::AssemblyPattern{APT_LinearForm,Float64,ON_CELLS}, FEB; ...)
# (reference: src/assemblypatterns/bilinearform.jl:384-400, src/assemblypatterns/linearform.jl:239-251)
# for the (FEType, operator, action) triples the library supports, and throws for anything else
# it is asked to handle explicitly -- not loading this file leaves the reference untouched.
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

# ---- codes of include/grmp.h ---------------------------------------------------------------
fecode(::Type{<:H1P1}) = 1
fecode(::Type{<:H1P2}) = 2
fecode(::Type{<:H1Pk{n,2,2}}) where {n} = 2     # same tables as H1P2 (DESIGN.md)
fecode(::Type{<:H1BR}) = 3
fecode(::Type{<:HDIVRT0}) = 4
fecode(::Type{<:HDIVBDM1}) = 5
fecode(::Type{<:L2P0}) = 6
opcode(::Type{Identity}) = 1
opcode(::Type{Gradient}) = 2
opcode(::Type{SymmetricGradient{1}}) = 3
opcode(::Type{Divergence}) = 4
opcode(::Type{ReconstructionIdentity{FER}}) where {FER<:HDIVRT0} = 5
opcode(::Type{ReconstructionIdentity{FER}}) where {FER<:HDIVBDM1} = 6
aptcode(::Type{GRMP.APT_BilinearForm}) = 0
aptcode(::Type{GRMP.APT_SymmetricBilinearForm}) = 1
aptcode(::Type{GRMP.APT_LumpedBilinearForm}) = 2

struct EvalTab
    nd_all::Int32
    ncomp::Int32
    refvals::Ptr{Float64}
    refderivs::Ptr{Float64}
end

# ---- handles (finalizers call the *_destroy entry points) ------------------------------------
mutable struct Ctx;   h::Ptr{Cvoid}; end
mutable struct DGrid; h::Ptr{Cvoid}; hasfaces::Bool; end
mutable struct DSpace; h::Ptr{Cvoid}; end
mutable struct DBlf;  h::Ptr{Cvoid}; nnz::Int64; colptr::Vector{Int64}; rowval::Vector{Int64}; end

const CTX = Ref{Union{Nothing,Ctx}}(nothing)
function context(device = 0)
    if CTX[] === nothing
        h = Ref{Ptr{Cvoid}}()
        check(ccall((:grmp_init, lib), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h))
        CTX[] = Ctx(h[])
    end
    return CTX[]
end

const GRIDS = IdDict{Any,DGrid}()
function device_grid(xgrid::ExtendableGrid{Float64,Int32}; faces = false)
    g = get!(GRIDS, xgrid) do
        coords = xgrid[Coordinates]; cn = xgrid[CellNodes]::Matrix{Int32}
        vol = xgrid[CellVolumes]; reg = Vector{Int32}(xgrid[CellRegions])
        h = Ref{Ptr{Cvoid}}()
        GC.@preserve coords cn vol reg check(ccall((:grmp_grid_create, lib), Cint,
            (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}, Ref{Ptr{Cvoid}}),
            context().h, size(coords, 1), size(coords, 2), coords, size(cn, 2), cn, vol, reg, h))
        d = DGrid(h[], false)
        finalizer(x -> ccall((:grmp_grid_destroy, lib), Cint, (Ptr{Cvoid},), x.h), d)
        d
    end
    if faces && !g.hasfaces
        cf = Matrix{Int32}(xgrid[CellFaces]); sg = Matrix{Int32}(xgrid[CellFaceSigns])
        ori = size(xgrid[Coordinates], 1) == 3 ? Matrix{Int32}(xgrid[CellFaceOrientations]) : Matrix{Int32}(undef, 0, 0)
        fn = xgrid[FaceNormals]; fv = xgrid[FaceVolumes]
        GC.@preserve cf sg ori fn fv check(ccall((:grmp_grid_set_faces, lib), Cint,
            (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
            g.h, length(fv), cf, sg, isempty(ori) ? C_NULL : pointer(ori), fn, fv))
        g.hasfaces = true
    end
    return g
end

const SPACES = IdDict{Any,DSpace}()
function device_space(FES::FESpace{Float64,Int32,FEType}) where {FEType}
    get!(SPACES, FES) do
        g = device_grid(FES.xgrid; faces = FEType <: Union{H1BR,HDIVRT0,HDIVBDM1})
        dofs = FES[CellDofs]
        colentries = dofs isa GRMP.SerialVariableTargetAdjacency ?
            Int32.(reshape(1:FES.ndofs, :, num_sources(FES.xgrid[CellNodes]))) : Matrix{Int32}(reshape(dofs.colentries, :, num_sources(dofs)))
        h = Ref{Ptr{Cvoid}}()
        GC.@preserve colentries check(ccall((:grmp_space_create, lib), Cint,
            (Ptr{Cvoid}, Cint, Cint, Int64, Cint, Ptr{Int32}, Ref{Ptr{Cvoid}}),
            g.h, fecode(FEType), get_ncomponents(FEType), FES.ndofs, size(colentries, 1), colentries, h))
        d = DSpace(h[])
        finalizer(x -> ccall((:grmp_space_destroy, lib), Cint, (Ptr{Cvoid},), x.h), d)
        d
    end
end

# tables straight out of the reference's FEEvaluator (ForwardDiff bits travel unchanged)
function evaltab(ev)   # ev::GRMP.SingleFEEvaluator
    nq = length(ev.xref)
    vals = isempty(ev.refbasisvals) ? Float64[] : permutedims(cat(ev.refbasisvals...; dims = 3), (2, 1, 3))[:]  # [comp, dof, i] -> memory [i][dof][comp]
    ders = ev.derivorder > 0 ? ev.refbasisderivvals[:] : Float64[]                                              # [row, j, i] column-major == [i][j][row]
    return vals, ders, EvalTab(size(ev.refbasisvals[1], 1), size(ev.refbasisvals[1], 2), pointer(vals), isempty(ders) ? C_NULL : pointer(ders))
end

const PATTERNS = IdDict{Any,DBlf}()

"""
    assemble!(A::FEMatrixBlock, AP; factor, skip_preps, ...)   (device version)

First call: prepare_assembly! (host, unchanged), grmp_blf_create + grmp_blf_symbolic + pattern download.
Every call: grmp_blf_numeric into a nzval buffer, then the block is installed / merged into
`A.entries` (single-block matrices: `A.entries.cscmatrix = SparseMatrixCSC(m, n, colptr, rowval, nzval)`;
otherwise `addblock!`-style merge through a temporary SparseMatrixCSC).
"""
function GRMP.assemble!(A::FEMatrixBlock{Float64,Int64,Float64,Int32}, AP::AssemblyPattern{APT,Float64,ON_CELLS}, FEB = [];
        factor = 1, factor_transpose = factor, skip_preps::Bool = false, fixed_arguments = nothing,
        transposed_assembly::Bool = false, transpose_copy = nothing) where {APT<:GRMP.APT_BilinearForm}
    length(FEB) == 0 || return invoke(GRMP.assemble!, Tuple{FEMatrixBlock,AssemblyPattern,Any}, A, AP, FEB; factor, skip_preps)  # 'next' row N4
    skip_preps || GRMP.prepare_assembly!(AP)
    d = get!(PATTERNS, AP) do
        e1 = GRMP.get_basisevaler(AP.AM, 1, 1); e2 = GRMP.get_basisevaler(AP.AM, 2, 1)
        v1, d1, t1 = evaltab(e1); v2, d2, t2 = evaltab(e2)
        w = GRMP.get_qweights(AP.AM)
        act, par = AP.action isa NoAction ? (0, Float64[]) : hooke_parameters(AP.action)   # Hooke tensors only; anything else throws
        regions = Vector{Int32}(AP.regions)
        h = Ref{Ptr{Cvoid}}()
        GC.@preserve v1 d1 v2 d2 w par regions check(ccall((:grmp_blf_create, lib), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float64}, Cint, Cint, Ptr{Int32}, Cint, Cint, Ptr{Float64}, Ref{EvalTab}, Ref{EvalTab}, Ref{Ptr{Cvoid}}),
            device_space(AP.FES[1]).h, device_space(AP.FES[2]).h, opcode(AP.operators[1]), opcode(AP.operators[2]), act,
            isempty(par) ? C_NULL : pointer(par), aptcode(APT), transposed_assembly, regions, length(regions), length(w), w, t1, t2, h))
        nnz = Ref{Int64}(0)
        check(ccall((:grmp_blf_symbolic, lib), Cint, (Ptr{Cvoid}, Float64, Ref{Int64}), h[], factor, nnz))
        colptr = Vector{Int64}(undef, size(A, 2) + 1); rowval = Vector{Int64}(undef, nnz[])
        check(ccall((:grmp_blf_get_pattern, lib), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), h[], colptr, rowval))
        b = DBlf(h[], nnz[], colptr, rowval)
        finalizer(x -> ccall((:grmp_blf_destroy, lib), Cint, (Ptr{Cvoid},), x.h), b)
        b
    end
    nzval = Vector{Float64}(undef, d.nnz)
    check(ccall((:grmp_blf_numeric, lib), Cint, (Ptr{Cvoid}, Float64, Ptr{Float64}), d.h, factor, nzval))
    B = SparseMatrixCSC(size(A, 1), size(A, 2), d.colptr, d.rowval, nzval)
    install_block!(A, B)
    if transpose_copy !== nothing
        cpt = Vector{Int64}(undef, size(A, 1) + 1); rvt = Vector{Int64}(undef, d.nnz); nzt = Vector{Float64}(undef, d.nnz)
        check(ccall((:grmp_blf_transpose_copy, lib), Cint, (Ptr{Cvoid}, Float64, Float64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
            d.h, factor, factor_transpose, cpt, rvt, nzt))
        install_block!(transpose_copy, SparseMatrixCSC(size(A, 2), size(A, 1), cpt, rvt, nzt))
    end
    AP.last_allocations = 0
    return nothing
end

"""
    assemble_from_host!(A, AP; factor)

Reassembly after the grid moved (same topology): one `grmp_blf_assemble_host` call uploads `Coordinates`, `CellVolumes`,
`CellNodes`, `CellDofs`, assembles on the frozen pattern and downloads `nzval` (uploads the kernels do not read overlap
the download).  `AP` must have been assembled once through `assemble!` above.
"""
function assemble_from_host!(A::FEMatrixBlock{Float64,Int64,Float64,Int32}, AP::AssemblyPattern; factor = 1)
    d = PATTERNS[AP]
    xgrid = AP.FES[1].xgrid
    coords = xgrid[Coordinates]; vol = xgrid[CellVolumes]; cn = xgrid[CellNodes]
    dofs1 = AP.FES[1][CellDofs].colentries
    dofs2 = AP.FES[2] === AP.FES[1] ? nothing : AP.FES[2][CellDofs].colentries
    nzval = Vector{Float64}(undef, d.nnz)
    GC.@preserve coords vol cn dofs1 dofs2 check(ccall((:grmp_blf_assemble_host, lib), Cint,
        (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}),
        d.h, factor, coords, vol, cn, dofs1, dofs2 === nothing ? C_NULL : pointer(dofs2), nzval))
    install_block!(A, SparseMatrixCSC(size(A, 1), size(A, 2), d.colptr, d.rowval, nzval))
    return nothing
end

# single-block FEMatrix with an empty target: adopt the CSC; otherwise merge (explicit zeros kept)
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

hooke_parameters(action) = error("only NoAction and the Hooke tensor actions of HookStiffnessOperator2D/3D are evaluated on the device; " *
                                 "construct the operator through GRMPCuda.HookStiffnessOperator2D to record (μ, λ)")

end # module
