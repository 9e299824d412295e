/* grmp.h -- C ABI of libgrmp_cuda, the B200-native replacement for the assembly hot path
 * of GradientRobustMultiPhysics.jl (BilinearForm / LinearForm AssemblyPatterns).
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * reference repository).  The Julia glue that binds these with `ccall` is shown in
 * INTEGRATION.md / julia/GRMPCuda.jl; the in-container binding is ctypes
 * (gradientrobustmultiphysics.jl_b200/_lib.py).
 *
 * Conventions
 *   - plain pointers and sizes only; all host arrays are owned by the caller and copied
 *     to the device inside the call; outputs are written into caller-allocated arrays.
 *   - indices on the wire are 1-based and laid out like Julia's column-major arrays
 *     (Coordinates dim x nnodes, CellNodes (dim+1) x ncells, CellDofs nd x ncells, ...):
 *     grid integers Int32 (ExtendableGrid{Float64,Int32}), matrix indices Int64
 *     (FEMatrix{Float64,Int64}, src/fematrix.jl:148-150).
 *   - every function returns 0 on success or a negative GRMP_E* code; the message is
 *     available from grmp_last_error() (thread-local).  Nothing throws or aborts.
 *   - calls are synchronous (the stream is synchronised before returning) unless the
 *     name ends in _async.
 *   - one context per process and device (one process per GPU; multi-GPU runs shard
 *     cells/columns over ranks on the host side, see DESIGN.md "multi-GPU").
 */
#ifndef GRMP_H
#define GRMP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRMP_OK 0
#define GRMP_EINVAL (-1)       /* bad argument */
#define GRMP_EUNSUPPORTED (-2) /* element / operator / action combination not on the ported path */
#define GRMP_ECUDA (-3)        /* CUDA runtime error (message carries cudaGetErrorString) */
#define GRMP_ENOMEM (-4)
#define GRMP_ESTATE (-5)       /* call order violated (e.g. numeric before symbolic) */

/* FEType codes (src/fedefs/{h1_p1,h1_p2,h1v_br,hdiv_rt0,hdiv_bdm1,l2_p0}.jl) */
enum { GRMP_FE_H1P1 = 1, GRMP_FE_H1P2 = 2, GRMP_FE_H1BR = 3, GRMP_FE_HDIVRT0 = 4, GRMP_FE_HDIVBDM1 = 5, GRMP_FE_L2P0 = 6 };
/* function operators (src/functionoperators.jl:13-153) */
enum { GRMP_OP_ID = 1, GRMP_OP_GRAD = 2, GRMP_OP_SYMGRAD = 3, GRMP_OP_DIV = 4, GRMP_OP_RECON_ID_RT0 = 5, GRMP_OP_RECON_ID_BDM1 = 6,
       GRMP_OP_NORMALFLUX = 7 /* Hdiv elements on boundary-face grids only: feevaluator_hdiv.jl:42-50 */ };
/* actions evaluated on the device (src/actions.jl:96-110 NoAction; src/pdeoperators.jl:265-270, 304-312 Hooke tensors) */
enum { GRMP_ACT_NONE = 0, GRMP_ACT_HOOKE2D = 1, GRMP_ACT_HOOKE3D = 2,
       GRMP_ACT_CONVECTION = 3 /* needs a fixed argument, see grmp_blf_set_fixed_argument */,
       GRMP_ACT_NEWTON_CONVECTION = 4 /* NonlinearForm, see grmp_blf_set_newton_argument */ };
/* assembly pattern types (src/assemblypatterns/bilinearform.jl:7-21) */
enum { GRMP_APT_BILINEARFORM = 0, GRMP_APT_SYMMETRIC = 1, GRMP_APT_LUMPED = 2 };
/* right-hand side data of a LinearForm (fdot_action, src/actions.jl:119-128) */
enum { GRMP_F_NONE = 0, GRMP_F_CONST = 1, GRMP_F_QP_TABLE = 2 };
/* numeric back ends of a bilinear form
 *   GENERIC   bit-exact two-phase path (reference operation and summation order, no FMA): the correctness path
 *   FAST      request the fastest owner-computes kernel that exists for the form (P2TET where applicable, else COLUMNS)
 *   P2TET     = what FAST resolves to for the metric form: ring-walk kernel of the 3D P2 Laplace stiffness matrix
 *   COLUMNS   owner-computes column kernels (one thread per matrix column, every non-zero written once, deterministic):
 *             H1 P1/P2/P0 (1..dim components), Bernardi-Raugel, RT0, BDM1; Identity/Gradient/SymmetricGradient+Hooke/Divergence
 *   ATOMIC    cell-parallel kernel, FP64 atomics through the per-cell local->nnz map (measured alternative, order not fixed)
 *   COLOURED  cell-parallel kernel under element colouring (deterministic; one launch per colour)
 *   AUTO      P2TET if applicable and the mesh passes the cancellation guard, else COLUMNS if a kernel exists, else GENERIC */
enum { GRMP_PATH_AUTO = 0, GRMP_PATH_GENERIC = 1, GRMP_PATH_FAST = 2, GRMP_PATH_P2TET = 2, GRMP_PATH_COLUMNS = 3, GRMP_PATH_ATOMIC = 4,
       GRMP_PATH_COLOURED = 5 };

typedef struct grmp_ctx grmp_ctx;
typedef struct grmp_grid grmp_grid;
typedef struct grmp_space grmp_space;
typedef struct grmp_blf grmp_blf;
typedef struct grmp_lf grmp_lf;

/* Reference-cell tables of one FEEvaluator (src/feevaluator.jl:34-138; reconstruction
 * constructor 142-217): in production they are taken from the Julia FEEvaluator
 * (`refbasisvals`, `refbasisderivvals`), so ForwardDiff's bits travel unchanged.
 *   refvals   [nq][nd_all][ncomp]            (refbasisvals[i][dof,comp]); for the
 *             ReconstructionIdentity operators: values of the Hdiv reconstruction basis
 *   refderivs [nq][edim][nd_all*ncomp]       (refbasisderivvals[dof+comp*nd_all, j, i]);
 *             NULL for operators that need no derivatives */
typedef struct grmp_evaltab {
  int32_t nd_all;
  int32_t ncomp;
  const double* refvals;
  const double* refderivs;
} grmp_evaltab;

/* timing / traffic counters of the last numeric call on a handle (the
 * library-side analogue of AP.last_allocations / SC.LHS_AssemblyTimes) */
typedef struct grmp_stats {
  double last_numeric_ms;   /* CUDA-event time of the last numeric phase */
  double last_symbolic_ms;
  int64_t nnz;
  int64_t ncontrib;         /* non-zero local contributions kept by the symbolic pass */
  int64_t kernel_launches;  /* kernels launched by the last numeric call */
  int32_t path;             /* GRMP_PATH_GENERIC or GRMP_PATH_FAST actually used */
  int32_t ntiles;
} grmp_stats;

const char* grmp_last_error(void);

/* context: selects the device, creates the stream */
int grmp_init(int device, grmp_ctx** out);
int grmp_finalize(grmp_ctx* ctx);
int grmp_device_synchronize(grmp_ctx* ctx);

/* Device-resident grid: what the assembly loop reads from `xgrid`
 * (bilinearform.jl:113-114: CellVolumes, CellRegions; feevaluator.jl:371-390: Coordinates,
 * CellNodes through the L2GTransformer).  cellregions may be NULL (all 1). */
int grmp_grid_create(grmp_ctx* ctx, int dim, int64_t nnodes, const double* coords, int64_t ncells,
                     const int32_t* cellnodes, const double* cellvolumes, const int32_t* cellregions,
                     grmp_grid** out);
/* The boundary faces as assembly items: AT = ON_BFACES (assemblypatterns.jl:400-440 reads BFaceNodes / BFaceVolumes /
 * BFaceRegions through GridComponent*4AssemblyType and FES[BFaceDofs] through Dofmap4AssemblyType, dofmaps.jl:45; call
 * sites: the best-approximation Dirichlet data of boundarydata.jl:297-347).  The returned grid has item dimension
 * xdim-1 (Edge1D / Triangle2D); spaces on it take FES[BFaceDofs] as `celldofs`.  Admitted: Identity of H1P1 / H1P2 (scalar
 * or vector valued) and NormalFlux of HDIVRT0 / HDIVBDM1 (face bases hdiv_rt0.jl:61-65, hdiv_bdm1.jl:74-79, 109-115; 1 / 2 / 3
 * dofs per face, evaluator tables with one component); everything else returns GRMP_EUNSUPPORTED and stays with the reference.
 * All forms on such a grid run on the bit-exact path. */
int grmp_grid_create_bfaces(grmp_ctx* ctx, int xdim, int64_t nnodes, const double* coords, int64_t nbfaces,
                            const int32_t* bfacenodes, const double* bfacevolumes, const int32_t* bfaceregions,
                            grmp_grid** out);
/* CellFaces / CellFaceSigns / CellFaceOrientations / FaceNormals / FaceVolumes
 * (hdiv_rt0.jl:106-116, hdiv_bdm1.jl coefficient + subset closures, h1v_br.jl:150-162,
 * 253-273, reconstructions.jl:27-30).  cellfaceorient may be NULL in 2D. */
int grmp_grid_set_faces(grmp_grid* grid, int64_t nfaces, const int32_t* cellfaces, const int32_t* cellfacesigns,
                        const int32_t* cellfaceorient, const double* facenormals, const double* facevolumes);
/* re-upload coordinates / volumes of an existing grid (moving meshes, the e2e bench step) */
int grmp_grid_update_geometry(grmp_grid* grid, const double* coords, const double* cellvolumes);
/* re-upload CellNodes of an existing grid (same sizes).  Invalidates the pattern of every form on the grid: numeric calls
 * return GRMP_ESTATE until grmp_blf_symbolic has run again. */
int grmp_grid_update_cells(grmp_grid* grid, const int32_t* cellnodes);
int grmp_grid_destroy(grmp_grid* grid);

/* FESpace + CellDofs (src/finiteelements.jl:42-49, src/dofmaps.jl:201-363): the dof map is
 * produced by the host (Julia: FES[CellDofs].colentries) */
int grmp_space_create(grmp_grid* grid, int fetype, int ncomp, int64_t ndofs, int nd_cell, const int32_t* celldofs,
                      grmp_space** out);
/* re-upload CellDofs of an existing space (same sizes); invalidates the pattern of every form on the space like grmp_grid_update_cells */
int grmp_space_update_dofs(grmp_space* space, const int32_t* celldofs);
int grmp_space_destroy(grmp_space* space);

/* AssemblyPattern{APT_BilinearForm...} + prepare_assembly! (bilinearform.jl:60-64,
 * assemblypatterns.jl:467-671): operators, action, regions, quadrature weights
 * (qf.w, nq of them) and the evaluator tables for both arguments.
 * space_row/op_row is the first ("ansatz", matrix row) argument: FES = [A.FESX, A.FESY]
 * (pdeoperators.jl:926-927).  transposed_assembly swaps row/column on output
 * (bilinearform.jl:353-357). */
int grmp_blf_create(grmp_space* space_row, grmp_space* space_col, int op_row, int op_col, int action,
                    const double* act_params, int apt, int transposed_assembly, const int32_t* regions, int nregions,
                    int nq, const double* qweights, const grmp_evaltab* tab_row, const grmp_evaltab* tab_col,
                    grmp_blf** out);
int grmp_blf_destroy(grmp_blf* blf);
/* choose the numeric back end before grmp_blf_symbolic (default GRMP_PATH_AUTO: the fast
 * owner-computes kernel where one exists for the form, else the generic two-phase path) */
int grmp_blf_set_path(grmp_blf* blf, int path);

/* Trilinear forms: assemble!(A, AP, FEB; fixed_arguments = [1]) with three FESpaces (src/assemblypatterns/bilinearform.jl:235-257): the
 * operator evaluation of the coefficient argument FEB[1] at every quadrature point is the first part of the action input.  On the device:
 * the Picard-linearised convection term of ConvectionOperator(a_from, a_operator, xdim, ncomponents; a_to = 1) (src/pdeoperators.jl:435-510),
 * action GRMP_ACT_CONVECTION: result[j] = sum_k a(x_q)[k] * (ansatz operator evaluation)[(j-1) xdim + k], i.e. ((a . grad) u, v).
 * space_a / op_a / tab_a describe FEB[1].FES and its operator, coeffs_host its entries (ndofs of space_a).  The table a(x_q) is evaluated on
 * the device in the reference's order (eval_febe!, feevaluator.jl:445-452).  The reference's pattern depends on the values: call
 * grmp_blf_symbolic afterwards, or pass keep_pattern = 1 to reassemble on the frozen pattern (the next Picard iteration). */
int grmp_blf_set_fixed_argument(grmp_blf* blf, grmp_space* space_a, int op_a, const grmp_evaltab* tab_a, const double* coeffs_host,
                                int keep_pattern);

/* NonlinearForm full_assemble!(A, b, AP, FEB) (src/assemblypatterns/nonlinearform.jl:44-245) for the Newton form of the convection term,
 * ConvectionOperator(a_from, a_operator, xdim, ncomponents; newton = true) (src/pdeoperators.jl:459-493): Jacobian matrix of
 * N(u) = ((a_operator(u) . ansatz_operator) u, test) at the current iterate and the right-hand side DN(u) u - N(u).
 * The form is created with grmp_blf_create(space_u, space_u, ansatz_operator, test_operator, GRMP_ACT_NEWTON_CONVECTION, NULL,
 * GRMP_APT_BILINEARFORM, transposed_assembly = 1, ...) and the rule prepare_assembly! picks for its THREE FESpaces; op_a / tab_a are
 * a_operator and its evaluator tables on space_u, coeffs_host the entries of the current iterate.  As in the reference the zero test of
 * _addnz is made on the unscaled local entry (nonlinearform.jl:216-221), so the pattern depends on the iterate: run grmp_blf_symbolic
 * afterwards, or pass keep_pattern = 1 for the next Newton step on the frozen pattern. */
int grmp_blf_set_newton_argument(grmp_blf* blf, int op_a, const grmp_evaltab* tab_a, const double* coeffs_host, int keep_pattern);
/* right-hand side of the last numeric call: b[dof + offset] += localb * itemfactor in cell order (nonlinearform.jl:226-233) */
int grmp_blf_newton_rhs(grmp_blf* blf, double* b_host, int64_t offset);

/* One-time symbolic pass on the GPU = what rawupdateindex! + flush! build on first
 * assembly (fematrix.jl:54-58, pdeoperators.jl:992): the pattern is the union of local
 * contributions with value != 0 (evaluated in the reference's operation order, no FMA),
 * which depends on `factor`.  Also builds the per-cell local->nnz map / gather lists. */
int grmp_blf_symbolic(grmp_blf* blf, double factor, int64_t* nnz_out);
/* SparseMatrixCSC{Float64,Int64} pattern: colptr[ncols+1], rowval[nnz], 1-based, rows ascending */
int grmp_blf_get_pattern(grmp_blf* blf, int64_t* colptr, int64_t* rowval);
/* Numeric assembly on the frozen pattern = assemble!(A, AP; factor, skip_preps = true) after
 * fill!(A, 0) (bilinearform.jl:92-380, solvers.jl:556).  nzval_host may be NULL (values stay
 * on the device; fetch later with grmp_blf_get_values). */
int grmp_blf_numeric(grmp_blf* blf, double factor, double* nzval_host);
int grmp_blf_get_values(grmp_blf* blf, double* nzval_host);
/* assemble!(A, AP; factor, skip_preps = true) with the grid still in HOST memory, one synchronous call (the end-to-end form
 * of bilinearform.jl:92-380 behind a Julia ccall): uploads Coordinates and CellVolumes, assembles on the frozen pattern and
 * downloads nzval. */
int grmp_blf_assemble_host(grmp_blf* blf, double factor, const double* coords, const double* cellvolumes,
                           const int32_t* cellnodes, const int32_t* celldofs_row, const int32_t* celldofs_col,
                           double* nzval_host);
/* Geometry (Coordinates, CellVolumes) is what may change on a frozen pattern.  cellnodes / celldofs_* may be NULL (trust the
 * frozen pattern, like skip_preps = true does); when given they are uploaded on the copy stream and COMPARED with the arrays of
 * the symbolic pass -- a difference returns GRMP_ESTATE (the pattern, the gather lists and the records are stale: call
 * grmp_grid_update_cells / grmp_space_update_dofs and grmp_blf_symbolic again).  nzval_host may be NULL (values stay resident). */
/* nsteps back-to-back numeric assemblies bracketed by ONE pair of CUDA events on the launching
 * stream (benchmarking / time loops that reassemble every step); total_ms receives the device time */
int grmp_blf_numeric_steps(grmp_blf* blf, double factor, int nsteps, double* total_ms);
/* Multi-GPU column ownership: only columns [0, ncols_owned) are assembled by the fast path (the
 * host numbers the columns a rank owns first; halo columns belong to another rank, DESIGN.md
 * "multi-GPU").  Call before grmp_blf_symbolic; ncols_owned < 0 restores "all columns". */
int grmp_blf_set_owned_columns(grmp_blf* blf, int64_t ncols_owned);
/* transpose_copy block (bilinearform.jl:358-364): CSC of the mirrored block with values
 * v * itemfactor / factor * factor_transpose * (-1), summed per entry in cell order.
 * Sizes: colptr_t[nrows+1], rowval_t[nnz], nzval_t[nnz]. Requires a prior numeric call. */
int grmp_blf_transpose_copy(grmp_blf* blf, double factor, double factor_transpose, int64_t* colptr_t,
                            int64_t* rowval_t, double* nzval_t);
int grmp_blf_stats(grmp_blf* blf, grmp_stats* out);
/* raw device pointer of nzval (for device-side consumers / benchmarks; owned by the library) */
int grmp_blf_device_values(grmp_blf* blf, void** dptr);

/* ---- what the reference does with the matrix right after assembly, without bringing it back to the host ------------------
 * Device-resident SparseMatrixCSC{Float64,Int64} of the last numeric call: the hand-off to a GPU solver in place of
 * `_LinearProblem(A.entries.cscmatrix, b.entries, SC)` (src/solvers.jl:655).  Pointers are device pointers owned by the
 * library, valid until the next symbolic call / destroy; colptr and rowval are 1-based Int64 like Julia's. */
typedef struct grmp_device_csc {
  int64_t nrows, ncols, nnz;
  const int64_t* colptr;   /* [ncols+1] */
  const int64_t* rowval;   /* [nnz] */
  const double* nzval;     /* [nnz] */
  int32_t device;
  int32_t reserved;
} grmp_device_csc;
int grmp_blf_device_csc(grmp_blf* blf, grmp_device_csc* out);
/* addblock_matmul!(a, B, b; factor, transposed) (src/fematrix.jl:402-473): a += B*b*factor, or a += B'*b*factor.  Host vectors
 * (a: nrows / ncols entries, updated in place).  Every a[i] receives its terms one at a time in the reference's order (columns
 * ascending, separate multiply and add), so the result is bit-identical to the reference's loop. */
int grmp_blf_matmul(grmp_blf* blf, const double* b_host, double* a_host, double factor, int transposed);
/* the same with device vectors (no copies; for device-resident solvers / time loops) */
int grmp_blf_matmul_device(grmp_blf* blf, const double* b_dev, double* a_dev, double factor, int transposed);
/* residual check of solve_direct! (src/solvers.jl:661-668): r = A*x - b, r[fixed_dofs] = 0; returns sum r_i^2 in *norm2.
 * fixed_dofs are 1-based and may be NULL; b_host may be NULL (r = A*x); r_host may be NULL (only the norm is wanted). */
int grmp_blf_residual(grmp_blf* blf, const double* x_host, const double* b_host, const int64_t* fixed_dofs, int64_t nfixed,
                      double* r_host, double* norm2);
/* apply_penalties!(A, fixed_dofs, penalty) (src/fematrix.jl:349-355): A[dof,dof] = penalty on the device-resident values.
 * The reference would INSERT a missing diagonal entry; a frozen pattern cannot grow, so the call fails with GRMP_EUNSUPPORTED
 * (and changes nothing else) if a fixed dof has no stored diagonal -- *nmissing (may be NULL) tells how many. */
int grmp_blf_apply_penalties(grmp_blf* blf, const int64_t* fixed_dofs, int64_t nfixed, double penalty, int64_t* nmissing);

/* AssemblyPattern{APT_LinearForm} (linearform.jl:29-33) with a single test-function argument */
int grmp_lf_create(grmp_space* space, int op, const int32_t* regions, int nregions, int nq, const double* qweights,
                   const grmp_evaltab* tab, grmp_lf** out);
int grmp_lf_destroy(grmp_lf* lf);
/* numeric back end of the linear form: GRMP_PATH_GENERIC (bit-exact two-phase path) or GRMP_PATH_COLUMNS (one thread per dof
 * gathers its cells' contributions in registers, cells ascending); GRMP_PATH_AUTO (default) takes COLUMNS where a kernel exists */
int grmp_lf_set_path(grmp_lf* lf, int path);
/* assemble!(b, AP; factor, offset) (linearform.jl:47-237): b[dof+offset] += contributions in
 * cell order.  fsrc = GRMP_F_NONE (no action: input = ones, 74-75), GRMP_F_CONST
 * (fdata[resultdim]) or GRMP_F_QP_TABLE (fdata[ncells][nq][resultdim], the host-evaluated
 * DataFunction).  b_host has length >= ndofs + offset and is updated in place. */
int grmp_lf_assemble(grmp_lf* lf, double factor, int fsrc, const double* fdata, double* b_host, int64_t offset);
/* assemble!(b, AP, FEB) of a LinearForm with ONE coefficient argument and NoAction (linearform.jl:130-178 with nFE = 2): the operator
 * evaluation of FEB[1] at the quadrature points takes the place of f, b[dof + offset] += int op_a(FEB[1]) . op(v_dof).  space_a / op_a /
 * tab_a describe FEB[1].FES and its operator (same result length as the test operator), coeffs_host its entries; the evaluation runs on
 * the device in the reference's order (eval_febe!, feevaluator.jl:445-452).  The rule of the linear form must be the one prepare_assembly!
 * picks for BOTH FESpaces (assemblypatterns.jl:559-565). */
int grmp_lf_assemble_feb(grmp_lf* lf, double factor, grmp_space* space_a, int op_a, const grmp_evaltab* tab_a, const double* coeffs_host,
                         double* b_host, int64_t offset);
int grmp_lf_stats(grmp_lf* lf, grmp_stats* out);

/* ---- ItemIntegrator with one argument (src/assemblypatterns/itemintegrator.jl:18-21, 160-360): the error norms every example ends
 * with.  The closed set of kernels evaluated on the device:
 *   GRMP_II_NONE     NoAction: b[j,item] += (operator evaluation of the FE function)[j] * w_i * |T|       (itemintegrator.jl:262-268)
 *   GRMP_II_L2NORM   L2NormIntegrator(ncomponents, operator)  (91-110): sum_j input[j]^2
 *   GRMP_II_L2ERROR  L2ErrorIntegrator(compare_data, operator; factor) (33-78): sum_j (compare_data(x)[j] - factor * input[j])^2, with
 *                    compare_data tabulated by the host at the quadrature points, data[ncells][nq][resultdim of the operator]
 * Other user Actions cannot cross a C ABI and stay on the reference path. */
enum { GRMP_II_NONE = 0, GRMP_II_L2NORM = 1, GRMP_II_L2ERROR = 2 };
typedef struct grmp_ii grmp_ii;
int grmp_ii_create(grmp_space* space, int op, int kind, const int32_t* regions, int nregions, int nq, const double* qweights,
                   const grmp_evaltab* tab, grmp_ii** out);
int grmp_ii_destroy(grmp_ii* ii);
/* length of the result per item: the operator's result length (NONE) or 1 */
int grmp_ii_resultdim(grmp_ii* ii, int* resultdim);
/* evaluate!(b, AP, FEB) (itemintegrator.jl:160-300): coeffs_host = the entries of the FEVectorBlock (ndofs of the space);
 * b_host [ncells][resultdim] (Julia: resultdim x nitems, column-major) is updated in place, b[j,item] += ..., bit-identical to the
 * reference's loop; may be NULL.  total_host[resultdim] (may be NULL) receives what evaluate(AP, FEB) (316-360) returns, the sum over
 * all items: the reference adds every (item, quadrature point) term to one running sum, the device adds the items' own sums in a
 * fixed tree order -- deterministic, equal to the reference to rounding (1e-12 relative for the norms, whose terms are >= 0). */
int grmp_ii_evaluate(grmp_ii* ii, const double* coeffs_host, double factor, const double* data_host, double* b_host, double* total_host);

#ifdef __cplusplus
}
#endif
#endif /* GRMP_H */
