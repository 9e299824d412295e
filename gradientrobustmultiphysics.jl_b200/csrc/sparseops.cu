// sparseops.cu -- matrix-vector products, residuals and Dirichlet penalties on the device-resident CSC matrix, so that the
// step after assemble! does not force a 1.9 GB download.
//
//   addblock_matmul!(a, B, b; factor, transposed)   src/fematrix.jl:402-473
//   mul!(residual, A, x); residual .-= b; residual[fixed_dofs] .= 0       src/solvers.jl:661-668
//   apply_penalties!(A, fixed_dofs, penalty)        src/fematrix.jl:349-355
//
// The reference walks the CSC column by column and adds `vals[r] * b[col] * factor` into a[row]; an entry a[row] therefore
// receives its terms in ascending column order.  The row-major view below lists the entries of every row in ascending column
// order, so one thread per row replays exactly that sequence (separate multiply and add) -> bit-identical results without
// atomics.  The transposed product reads the CSC directly (one thread per column, rows ascending).
#include <cub/cub.cuh>

#include "sparseops.cuh"

namespace grmp {

namespace {

inline unsigned nblk(i64 n, int t = 256) { return (unsigned)((n + t - 1) / t); }

__global__ void csr_fill(const i64* rowval_t, const i64* colptr_t, i64 nrows, i64 nnz, i32* col, i64* rowptr) {
  const i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (k < nnz) col[k] = (i32)(rowval_t[k] - 1);
  if (k <= nrows) rowptr[k] = colptr_t[k] - 1;
}

// a[row] += sum_k (nzval[slot[k]] * b[col[k]]) * factor, k ascending (= columns ascending)
__global__ void __launch_bounds__(256) matmul_rows(const i64* __restrict__ rowptr, const i32* __restrict__ col, const i32* __restrict__ slot,
                                                   const double* __restrict__ nzval, const double* __restrict__ b, double* __restrict__ a,
                                                   i64 nrows, double factor) {
  const i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (i >= nrows) return;
  double acc = a[i];
  const i64 k1 = rowptr[i + 1];
  for (i64 k = rowptr[i]; k < k1; k++) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(nzval[slot[k]], b[col[k]]), factor));
  a[i] = acc;
}
// a[col] += sum_r (nzval[r] * b[row(r)]) * factor, r ascending
__global__ void __launch_bounds__(256) matmul_cols(const i64* __restrict__ colptr, const i64* __restrict__ rowval, const double* __restrict__ nzval,
                                                   const double* __restrict__ b, double* __restrict__ a, i64 ncols, double factor) {
  const i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (j >= ncols) return;
  double acc = a[j];
  const i64 k1 = colptr[j + 1] - 1;
  for (i64 k = colptr[j] - 1; k < k1; k++) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(nzval[k], b[rowval[k] - 1]), factor));
  a[j] = acc;
}

__global__ void penalties_kernel(const i64* colptr, const i64* rowval, i64 ncols, i64 nrows, double* nzval, const i64* fixed, i64 nfixed,
                                 double penalty, unsigned long long* missing) {
  const i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t >= nfixed) return;
  const i64 d = fixed[t];                      // 1-based
  if (d < 1 || d > ncols || d > nrows) { atomicAdd(missing, 1ull); return; }
  i64 lo = colptr[d - 1] - 1, hi = colptr[d] - 1;
  const i64 end = hi;
  while (lo < hi) { const i64 mid = (lo + hi) >> 1; if (rowval[mid] < d) lo = mid + 1; else hi = mid; }
  if (lo < end && rowval[lo] == d) nzval[lo] = penalty;
  else atomicAdd(missing, 1ull);
}

__global__ void sub_kernel(double* r, const double* b, i64 n) {
  const i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (i < n) r[i] = __dadd_rn(r[i], -b[i]);
}
__global__ void zero_fixed(double* r, i64 n, const i64* fixed, i64 nfixed) {
  const i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t < nfixed && fixed[t] >= 1 && fixed[t] <= n) r[fixed[t] - 1] = 0.0;
}
__global__ void square_kernel(const double* r, i64 n, double* out) {
  const i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (i < n) out[i] = r[i] * r[i];
}

}  // namespace

int build_csr_view(cudaStream_t s, const Pattern& pat, CsrView* out) {
  out->built = false;
  if (pat.nnz >= (i64)0x7fffffff) return fail(GRMP_EUNSUPPORTED, "row view: more than 2^31-1 non-zeros on one device");
  DevBuf<i64> colptr_t, rowval_t;
  GRMP_TRY(build_transposed(s, pat, colptr_t, rowval_t, out->slot));
  GRMP_TRY(out->rowptr.alloc(pat.nrows + 1));
  GRMP_TRY(out->col.alloc(std::max<i64>(pat.nnz, 1)));
  csr_fill<<<nblk(std::max(pat.nnz, pat.nrows + 1)), 256, 0, s>>>(rowval_t.p, colptr_t.p, pat.nrows, pat.nnz, out->col.p, out->rowptr.p);
  GRMP_CUDA(cudaGetLastError());
  GRMP_CUDA(cudaStreamSynchronize(s));
  out->built = true;
  return GRMP_OK;
}

int launch_matmul(cudaStream_t s, const Pattern& pat, const CsrView& csr, const double* nzval, const double* b, double* a, double factor,
                  int transposed) {
  if (transposed) {
    if (pat.ncols > 0) matmul_cols<<<nblk(pat.ncols), 256, 0, s>>>(pat.colptr.p, pat.rowval.p, nzval, b, a, pat.ncols, factor);
  } else {
    if (!csr.built) return fail(GRMP_ESTATE, "row view not built");
    if (pat.nrows > 0) matmul_rows<<<nblk(pat.nrows), 256, 0, s>>>(csr.rowptr.p, csr.col.p, csr.slot.p, nzval, b, a, pat.nrows, factor);
  }
  GRMP_CUDA(cudaGetLastError());
  return GRMP_OK;
}

int launch_penalties(cudaStream_t s, const Pattern& pat, double* nzval, const i64* fixed, i64 nfixed, double penalty, i64* missing_dev) {
  GRMP_CUDA(cudaMemsetAsync(missing_dev, 0, 8, s));
  if (nfixed > 0) {
    penalties_kernel<<<nblk(nfixed), 256, 0, s>>>(pat.colptr.p, pat.rowval.p, pat.ncols, pat.nrows, nzval, fixed, nfixed, penalty,
                                                 reinterpret_cast<unsigned long long*>(missing_dev));
    GRMP_CUDA(cudaGetLastError());
  }
  return GRMP_OK;
}

int launch_residual_finish(cudaStream_t s, double* r, const double* b, i64 n, const i64* fixed, i64 nfixed, double* norm2_dev) {
  if (n > 0 && b) sub_kernel<<<nblk(n), 256, 0, s>>>(r, b, n);
  if (nfixed > 0) zero_fixed<<<nblk(nfixed), 256, 0, s>>>(r, n, fixed, nfixed);
  GRMP_CUDA(cudaGetLastError());
  if (norm2_dev) {
    DevBuf<double> sq; DevBuf<unsigned char> temp;
    GRMP_TRY(sq.alloc(std::max<i64>(n, 1)));
    GRMP_CUDA(cudaMemsetAsync(sq.p, 0, 8, s));
    if (n > 0) square_kernel<<<nblk(n), 256, 0, s>>>(r, n, sq.p);
    size_t tb = 0;
    GRMP_CUDA(cub::DeviceReduce::Sum(nullptr, tb, sq.p, norm2_dev, std::max<i64>(n, 1), s));
    GRMP_TRY(temp.alloc(tb));
    GRMP_CUDA(cub::DeviceReduce::Sum(temp.p, tb, sq.p, norm2_dev, std::max<i64>(n, 1), s));
    GRMP_CUDA(cudaStreamSynchronize(s));
  }
  return GRMP_OK;
}

int device_sum(cudaStream_t s, const double* x, i64 n, double* out_dev, DevBuf<unsigned char>* tmp) {
  if (n <= 0) { GRMP_CUDA(cudaMemsetAsync(out_dev, 0, 8, s)); return GRMP_OK; }
  size_t tb = 0;
  GRMP_CUDA(cub::DeviceReduce::Sum(nullptr, tb, x, out_dev, n, s));
  if (tb > tmp->n) GRMP_TRY(tmp->alloc(std::max<size_t>(tb, 16)));
  GRMP_CUDA(cub::DeviceReduce::Sum(tmp->p, tb, x, out_dev, n, s));
  return GRMP_OK;
}

}  // namespace grmp
