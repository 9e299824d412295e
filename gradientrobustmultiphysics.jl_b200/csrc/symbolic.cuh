// symbolic.cuh -- one-time GPU symbolic pass + ordered gather (generic numeric back end)
#pragma once
#include "common.cuh"

namespace grmp {

// Result of the symbolic pass of one bilinear form: the CSC pattern that
// rawupdateindex! + flush! would build (union of non-zero local contributions, rows
// ascending per column, explicit zeros from later cancellation kept), plus the gather
// lists that replay the reference's summation order (cells ascending).
struct Pattern {
  i64 nrows = 0, ncols = 0, nnz = 0, ncontrib = 0;
  DevBuf<i64> colptr;     // [ncols+1] 1-based
  DevBuf<i64> rowval;     // [nnz]     1-based
  DevBuf<i32> colidx;     // [nnz]     0-based column of every slot
  DevBuf<i64> segptr;     // [nnz+1]   contributions of slot s: gsrc[segptr[s] .. segptr[s+1])
  DevBuf<u32> gsrc;       // [ncontrib] index into the element-matrix buffer lbuf ([entry][cell])
  DevBuf<i32> slotmap;    // [ncells*nd1*nd2] per-cell local -> nnz map (-1: masked out); optional
};

// keys: [ntot] u64 (col*nrows+row, ~0 = masked), cell-major; consumed (sorted in place / freed)
int build_pattern(cudaStream_t s, DevBuf<u64>& keys, i64 ntot, i64 nrows, i64 ncols, i64 ncells, int nd1, int nd2,
                  bool symmetric, bool want_slotmap, Pattern* out);

// nzval[s] = sum_k lbuf[gsrc[k]] in contribution order, starting from 0.0
int launch_gather(cudaStream_t s, const Pattern& pat, const double* lbuf, double* nzval);
// transposed-copy values: sum_k ((lbuf[gsrc[k]] / factor * factor_transpose) * -1)
int launch_gather_transposed(cudaStream_t s, const Pattern& pat, const double* lbuf, double factor, double factor_transpose,
                             double* tvals);
// CSC of the transposed pattern: colptr_t [nrows+1], rowval_t [nnz] (1-based) and perm[pos] = source slot
int build_transposed(cudaStream_t s, const Pattern& pat, DevBuf<i64>& colptr_t, DevBuf<i64>& rowval_t, DevBuf<i32>& perm);
int launch_permute(cudaStream_t s, const double* src, const i32* perm, i64 n, double* dst);
int launch_scale(cudaStream_t s, const double* src, i64 n, double alpha, double* dst);

// LinearForm: dof -> (cell, local dof) lists in cell order
struct DofGather {
  i64 ndofs = 0, ncontrib = 0;
  DevBuf<i64> segptr;   // [ndofs+1]
  DevBuf<u32> gsrc;     // [ncontrib] index into lbuf ([d][cell])
  DevBuf<u32> gcell;    // [ncontrib] cell of the contribution (region filter)
};
int build_dofgather(cudaStream_t s, const i32* celldofs, i64 ncells, int nd, i64 ndofs, DofGather* out);
int launch_lf_gather(cudaStream_t s, const DofGather& dg, const double* lbuf, const unsigned char* active, double* b);

}  // namespace grmp
