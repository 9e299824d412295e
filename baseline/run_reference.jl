# run_reference.jl -- runs the REAL reference (needs Julia >= 1.6 with GradientRobustMultiPhysics v0.12,
# ExtendableGrids >= 0.9.16, ExtendableSparse >= 1.2; NOT available in the build container).
# It (1) times assemble_operator! for the benchmark configuration exactly like bench.py times the
# port (first assembly = pattern + values; reassembly after fill!(A,0) with skip_preps = true) and
# (2) dumps grid arrays + colptr/rowval/nzval so that oracle and libgrmp_cuda can be checked on
# reference-generated inputs, closing the "parity unpinned" gap of DESIGN.md 6.
#
#   JULIA_NUM_THREADS=1 julia baseline/run_reference.jl 4 out_dir
using GradientRobustMultiPhysics, ExtendableGrids, ExtendableSparse, SparseArrays, DelimitedFiles

level = length(ARGS) > 0 ? parse(Int, ARGS[1]) : 4
outdir = length(ARGS) > 1 ? ARGS[2] : "reference_dump"
mkpath(outdir)

xgrid = uniform_refine(grid_unitcube(Tetrahedron3D), level)
FES = FESpace{H1P2{1,3}}(xgrid)
A = FEMatrix{Float64}(FES)
O = LaplaceOperator(1.0)

assemble_operator!(A[1,1], O)                       # warm-up / compilation
A = FEMatrix{Float64}(FES)
t_first = @elapsed assemble_operator!(A[1,1], O)    # LNK insertion + flush!
P = GradientRobustMultiPhysics.create_assembly_pattern(O, A[1,1], nothing)
assemble_operator!(A[1,1], O; Pattern = P)          # prepares P
fill!(A[1,1], 0)
t_re = @elapsed assemble_operator!(A[1,1], O; Pattern = P, skip_preps = true)
csc = A.entries.cscmatrix
println("threads = ", Threads.nthreads(), "  level = ", level, "  ncells = ", num_sources(xgrid[CellNodes]),
        "  ndofs = ", FES.ndofs, "  nnz = ", nnz(csc))
println("first assembly  ", t_first, " s  ", nnz(csc) / t_first, " nnz/s")
println("reassembly      ", t_re, " s  ", nnz(csc) / t_re, " nnz/s")

dump(name, a) = open(io -> write(io, a), joinpath(outdir, name), "w")
dump("coords.f64", xgrid[Coordinates]); dump("cellnodes.i32", Matrix{Int32}(xgrid[CellNodes]))
dump("cellvolumes.f64", xgrid[CellVolumes]); dump("celldofs.i32", Int32.(FES[CellDofs].colentries))
dump("celledges.i32", Matrix{Int32}(xgrid[CellEdges])); dump("cellfaces.i32", Matrix{Int32}(xgrid[CellFaces]))
dump("cellfacesigns.i32", Matrix{Int32}(xgrid[CellFaceSigns])); dump("cellfaceorientations.i32", Matrix{Int32}(xgrid[CellFaceOrientations]))
dump("facenormals.f64", xgrid[FaceNormals]); dump("facevolumes.f64", xgrid[FaceVolumes])
dump("colptr.i64", csc.colptr); dump("rowval.i64", csc.rowval); dump("nzval.f64", csc.nzval)
