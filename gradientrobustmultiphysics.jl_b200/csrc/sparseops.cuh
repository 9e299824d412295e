// sparseops.cuh -- what the reference does with the assembled matrix right after assembly, on the device-resident CSC:
// addblock_matmul! / mul! (fematrix.jl:402-473, solvers.jl:661-668), apply_penalties! (fematrix.jl:349-355): see sparseops.cu
#pragma once
#include "common.cuh"
#include "symbolic.cuh"

namespace grmp {

// row-major view of the CSC pattern (built once per pattern): rows of A = columns of A^T
struct CsrView {
  bool built = false;
  DevBuf<i64> rowptr;     // [nrows+1], 0-based
  DevBuf<i32> col;        // [nnz] column of every entry, ascending inside a row
  DevBuf<i32> slot;       // [nnz] position of the entry in the CSC arrays
};
int build_csr_view(cudaStream_t s, const Pattern& pat, CsrView* out);
// a += B * b * factor (transposed = 0) or a += B^T * b * factor (transposed = 1), one term at a time in the reference's order
// (columns ascending, rows ascending inside a column; no FMA) -> bit-identical to addblock_matmul!
int launch_matmul(cudaStream_t s, const Pattern& pat, const CsrView& csr, const double* nzval, const double* b, double* a, double factor,
                  int transposed);
// A[dof, dof] = penalty for every fixed dof (1-based); *missing = number of fixed dofs whose diagonal is not in the pattern
int launch_penalties(cudaStream_t s, const Pattern& pat, double* nzval, const i64* fixed_dofs_dev, i64 nfixed, double penalty, i64* missing_dev);
// r = r - b, r[fixed] = 0, *norm2 = sum r^2 (deterministic two-stage reduction)
int launch_residual_finish(cudaStream_t s, double* r, const double* b, i64 n, const i64* fixed_dofs_dev, i64 nfixed, double* norm2_dev);

// *out_dev = sum x[0..n) (cub::DeviceReduce: fixed tree for a given n -> deterministic); tmp is reused between calls
int device_sum(cudaStream_t s, const double* x, i64 n, double* out_dev, DevBuf<unsigned char>* tmp);

}  // namespace grmp
