// symbolic.cu -- GPU symbolic pass replacing rawupdateindex!/flush! pattern construction
// (ExtendableSparse LNK -> CSC; call sites src/fematrix.jl:54-65, src/pdeoperators.jl:992)
// and the ordered gather that replays the reference's per-entry summation order.
#include <cub/cub.cuh>

#include "symbolic.cuh"

namespace grmp {

namespace {

__global__ void iota_u32(u32* a, i64 n) {
  i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (i < n) a[i] = (u32)i;
}

// first index whose (masked) key equals the sentinel
__global__ void find_nvalid(const u64* keys, i64 n, u64 sentinel, i64* out) {
  i64 lo = 0, hi = n;
  while (lo < hi) {
    i64 mid = (lo + hi) >> 1;
    if (keys[mid] >= sentinel) hi = mid; else lo = mid + 1;
  }
  *out = lo;
}

__global__ void head_flags(const u64* keys, i64 n, u32* flags) {
  i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (k < n) flags[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1u : 0u;
}

__global__ void fill_slots(const u64* keys, const u32* ids, const u32* incl, i64 nvalid, i64 nrows, i64 ncells, int nd1, int nd2,
                           int symmetric, i64* rowval, i32* colidx, i64* segptr, u32* gsrc, i32* slotmap) {
  i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (k >= nvalid) return;
  const u32 slot = incl[k] - 1;
  const u64 key = keys[k];
  if (k == 0 || keys[k - 1] != key) {
    rowval[slot] = (i64)(key % (u64)nrows) + 1;
    colidx[slot] = (i32)(key / (u64)nrows);
    segptr[slot] = k;
  }
  const int nloc = nd1 * nd2;
  const u32 id = ids[k];
  const i64 cell = id / nloc;
  int e = id % nloc;
  if (symmetric) {
    int di = e / nd2, dj = e % nd2;
    if (dj < di) e = dj * nd2 + di;   // lower entries read the mirrored upper value (bilinearform.jl:330-341)
  }
  gsrc[k] = (u32)((i64)e * ncells + cell);
  if (slotmap) slotmap[id] = (i32)slot;
}

__global__ void colptr_from_colidx(const i32* colidx, i64 nnz, i64 ncols, i64* colptr) {
  i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (j > ncols) return;
  i64 lo = 0, hi = nnz;   // first slot with colidx >= j
  while (lo < hi) {
    i64 mid = (lo + hi) >> 1;
    if (colidx[mid] >= j) hi = mid; else lo = mid + 1;
  }
  colptr[j] = lo + 1;
}

__global__ void fill_i32(i32* a, i64 n, i32 v) {
  i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}

__global__ void gather_kernel(const i64* __restrict__ segptr, const u32* __restrict__ gsrc, const double* __restrict__ lbuf, i64 nnz,
                              double* __restrict__ nzval) {
  i64 s = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (s >= nnz) return;
  double sum = 0.0;
  for (i64 k = segptr[s]; k < segptr[s + 1]; k++) sum += lbuf[gsrc[k]];
  nzval[s] = sum;
}

__global__ void gather_transposed_kernel(const i64* __restrict__ segptr, const u32* __restrict__ gsrc, const double* __restrict__ lbuf,
                                         i64 nnz, double factor, double ft, double* __restrict__ tv) {
  i64 s = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (s >= nnz) return;
  double sum = 0.0;
  for (i64 k = segptr[s]; k < segptr[s + 1]; k++) sum += (lbuf[gsrc[k]] / factor * ft) * -1.0;
  tv[s] = sum;
}

__global__ void transposed_keys(const i64* rowval, const i32* colidx, i64 nnz, i64 ncols, u64* keys, u32* ids) {
  i64 s = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (s >= nnz) return;
  keys[s] = (u64)(rowval[s] - 1) * (u64)ncols + (u64)colidx[s];
  ids[s] = (u32)s;
}
__global__ void transposed_fill(const u64* keys, const u32* ids, i64 nnz, i64 ncols, i64* rowval_t, i32* colidx_t, i32* perm) {
  i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (k >= nnz) return;
  rowval_t[k] = (i64)(keys[k] % (u64)ncols) + 1;
  colidx_t[k] = (i32)(keys[k] / (u64)ncols);
  perm[k] = (i32)ids[k];
}
__global__ void permute_kernel(const double* src, const i32* perm, i64 n, double* dst) {
  i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (k < n) dst[k] = src[perm[k]];
}

__global__ void dof_keys(const i32* celldofs, i64 n, u32* keys, u32* ids) {
  i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (i < n) { keys[i] = (u32)(celldofs[i] - 1); ids[i] = (u32)i; }
}
__global__ void dof_fill(const u32* ids, i64 n, i64 ncells, int nd, u32* gsrc, u32* gcell) {
  i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (k >= n) return;
  u32 id = ids[k];
  i64 cell = id / nd; int d = id % nd;
  gsrc[k] = (u32)((i64)d * ncells + cell);
  gcell[k] = (u32)cell;
}
__global__ void segptr_from_sorted_u32(const u32* keys, i64 n, i64 nseg, i64* segptr) {
  i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (j > nseg) return;
  i64 lo = 0, hi = n;
  while (lo < hi) {
    i64 mid = (lo + hi) >> 1;
    if ((i64)keys[mid] >= j) hi = mid; else lo = mid + 1;
  }
  segptr[j] = lo;
}
__global__ void lf_gather_kernel(const i64* __restrict__ segptr, const u32* __restrict__ gsrc, const u32* __restrict__ gcell,
                                 const double* __restrict__ lbuf, const unsigned char* __restrict__ active, i64 ndofs, double* b) {
  i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (j >= ndofs) return;
  double sum = b[j];
  for (i64 k = segptr[j]; k < segptr[j + 1]; k++)
    if (active[gcell[k]]) sum += lbuf[gsrc[k]];   // b[dof] += localb * itemfactor, cells ascending (linearform.jl:216-220)
  b[j] = sum;
}

inline unsigned nblk(i64 n, int t = 256) { return (unsigned)((n + t - 1) / t); }

int bits_for(u64 maxval) {
  int b = 1;
  while (b < 64 && (maxval >> b) != 0) b++;
  return b;
}

}  // namespace

int build_pattern(cudaStream_t s, DevBuf<u64>& keys, i64 ntot, i64 nrows, i64 ncols, i64 ncells, int nd1, int nd2, bool symmetric,
                  bool want_slotmap, Pattern* out) {
  out->nrows = nrows; out->ncols = ncols; out->nnz = 0; out->ncontrib = 0;
  GRMP_TRY(out->colptr.alloc(ncols + 1));
  if (ntot >= (i64)0xffffffffll) return fail(GRMP_EUNSUPPORTED, "more than 2^32 local contributions on one device");
  if ((double)nrows * (double)ncols >= 9.0e18) return fail(GRMP_EUNSUPPORTED, "matrix too large for 64-bit keys");
  if (ntot == 0) {
    colptr_from_colidx<<<nblk(ncols + 1), 256, 0, s>>>(nullptr, 0, ncols, out->colptr.p);
    GRMP_CUDA(cudaGetLastError());
    return GRMP_OK;
  }
  // sort (key, contribution id); radix sort is stable, so equal keys stay in cell order
  DevBuf<u64> keys2; DevBuf<u32> ids, ids2; DevBuf<unsigned char> temp;
  GRMP_TRY(keys2.alloc(ntot)); GRMP_TRY(ids.alloc(ntot)); GRMP_TRY(ids2.alloc(ntot));
  iota_u32<<<nblk(ntot), 256, 0, s>>>(ids.p, ntot);
  const int end_bit = bits_for((u64)nrows * (u64)ncols) + 1;   // sentinel (all ones) sorts last
  cub::DoubleBuffer<u64> dk(keys.p, keys2.p);
  cub::DoubleBuffer<u32> dv(ids.p, ids2.p);
  size_t tb = 0;
  GRMP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, ntot, 0, end_bit, s));
  GRMP_TRY(temp.alloc(tb));
  GRMP_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tb, dk, dv, ntot, 0, end_bit, s));
  const u64* sk = dk.Current();
  const u32* sv = dv.Current();
  // number of unmasked contributions
  DevBuf<i64> scalar; GRMP_TRY(scalar.alloc(1));
  const u64 sentinel_masked = ~0ull;
  find_nvalid<<<1, 1, 0, s>>>(sk, ntot, sentinel_masked, scalar.p);
  i64 nvalid = 0;
  GRMP_CUDA(cudaMemcpyAsync(&nvalid, scalar.p, 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  out->ncontrib = nvalid;
  if (want_slotmap) {
    GRMP_TRY(out->slotmap.alloc(ntot));
    fill_i32<<<nblk(ntot), 256, 0, s>>>(out->slotmap.p, ntot, -1);
  }
  if (nvalid == 0) {
    colptr_from_colidx<<<nblk(ncols + 1), 256, 0, s>>>(nullptr, 0, ncols, out->colptr.p);
    GRMP_CUDA(cudaGetLastError());
    return GRMP_OK;
  }
  // unique keys -> slots
  DevBuf<u32> flags; GRMP_TRY(flags.alloc(nvalid));
  head_flags<<<nblk(nvalid), 256, 0, s>>>(sk, nvalid, flags.p);
  size_t tb2 = 0;
  GRMP_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb2, flags.p, flags.p, nvalid, s));
  if (tb2 > temp.n) GRMP_TRY(temp.alloc(tb2));
  GRMP_CUDA(cub::DeviceScan::InclusiveSum(temp.p, tb2, flags.p, flags.p, nvalid, s));
  u32 nnz32 = 0;
  GRMP_CUDA(cudaMemcpyAsync(&nnz32, flags.p + (nvalid - 1), 4, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  const i64 nnz = nnz32;
  out->nnz = nnz;
  GRMP_TRY(out->rowval.alloc(nnz)); GRMP_TRY(out->colidx.alloc(nnz)); GRMP_TRY(out->segptr.alloc(nnz + 1));
  GRMP_TRY(out->gsrc.alloc(nvalid));
  if ((double)nd1 * nd2 * (double)ncells >= 4294967295.0) return fail(GRMP_EUNSUPPORTED, "element-matrix buffer exceeds 2^32 entries");
  fill_slots<<<nblk(nvalid), 256, 0, s>>>(sk, sv, flags.p, nvalid, nrows, ncells, nd1, nd2, symmetric ? 1 : 0, out->rowval.p,
                                         out->colidx.p, out->segptr.p, out->gsrc.p, want_slotmap ? out->slotmap.p : nullptr);
  GRMP_CUDA(cudaMemcpyAsync(out->segptr.p + nnz, &nvalid, 8, cudaMemcpyHostToDevice, s));
  colptr_from_colidx<<<nblk(ncols + 1), 256, 0, s>>>(out->colidx.p, nnz, ncols, out->colptr.p);
  GRMP_CUDA(cudaGetLastError());
  GRMP_CUDA(cudaStreamSynchronize(s));
  keys.release();
  return GRMP_OK;
}


int launch_gather(cudaStream_t s, const Pattern& pat, const double* lbuf, double* nzval) {
  if (pat.nnz == 0) return GRMP_OK;
  gather_kernel<<<nblk(pat.nnz), 256, 0, s>>>(pat.segptr.p, pat.gsrc.p, lbuf, pat.nnz, nzval);
  GRMP_CUDA(cudaGetLastError());
  return GRMP_OK;
}

int launch_gather_transposed(cudaStream_t s, const Pattern& pat, const double* lbuf, double factor, double ft, double* tv) {
  if (pat.nnz == 0) return GRMP_OK;
  gather_transposed_kernel<<<nblk(pat.nnz), 256, 0, s>>>(pat.segptr.p, pat.gsrc.p, lbuf, pat.nnz, factor, ft, tv);
  GRMP_CUDA(cudaGetLastError());
  return GRMP_OK;
}

int build_transposed(cudaStream_t s, const Pattern& pat, DevBuf<i64>& colptr_t, DevBuf<i64>& rowval_t, DevBuf<i32>& perm) {
  const i64 nnz = pat.nnz;
  GRMP_TRY(colptr_t.alloc(pat.nrows + 1));
  GRMP_TRY(rowval_t.alloc(nnz)); GRMP_TRY(perm.alloc(nnz));
  if (nnz == 0) {
    colptr_from_colidx<<<nblk(pat.nrows + 1), 256, 0, s>>>(nullptr, 0, pat.nrows, colptr_t.p);
    GRMP_CUDA(cudaGetLastError());
    return GRMP_OK;
  }
  DevBuf<u64> k1, k2; DevBuf<u32> v1, v2; DevBuf<unsigned char> temp; DevBuf<i32> colidx_t;
  GRMP_TRY(k1.alloc(nnz)); GRMP_TRY(k2.alloc(nnz)); GRMP_TRY(v1.alloc(nnz)); GRMP_TRY(v2.alloc(nnz)); GRMP_TRY(colidx_t.alloc(nnz));
  transposed_keys<<<nblk(nnz), 256, 0, s>>>(pat.rowval.p, pat.colidx.p, nnz, pat.ncols, k1.p, v1.p);
  cub::DoubleBuffer<u64> dk(k1.p, k2.p);
  cub::DoubleBuffer<u32> dv(v1.p, v2.p);
  size_t tb = 0;
  const int end_bit = bits_for((u64)pat.nrows * (u64)pat.ncols) + 1;
  GRMP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, nnz, 0, end_bit, s));
  GRMP_TRY(temp.alloc(tb));
  GRMP_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tb, dk, dv, nnz, 0, end_bit, s));
  transposed_fill<<<nblk(nnz), 256, 0, s>>>(dk.Current(), dv.Current(), nnz, pat.ncols, rowval_t.p, colidx_t.p, perm.p);
  colptr_from_colidx<<<nblk(pat.nrows + 1), 256, 0, s>>>(colidx_t.p, nnz, pat.nrows, colptr_t.p);
  GRMP_CUDA(cudaGetLastError());
  GRMP_CUDA(cudaStreamSynchronize(s));
  return GRMP_OK;
}

__global__ void scale_kernel(const double* src, i64 n, double alpha, double* dst) {
  i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (k < n) dst[k] = src[k] * alpha;
}
int launch_scale(cudaStream_t s, const double* src, i64 n, double alpha, double* dst) {
  if (n == 0) return GRMP_OK;
  scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, n, alpha, dst);
  GRMP_CUDA(cudaGetLastError());
  return GRMP_OK;
}

int launch_permute(cudaStream_t s, const double* src, const i32* perm, i64 n, double* dst) {
  if (n == 0) return GRMP_OK;
  permute_kernel<<<nblk(n), 256, 0, s>>>(src, perm, n, dst);
  GRMP_CUDA(cudaGetLastError());
  return GRMP_OK;
}

int build_dofgather(cudaStream_t s, const i32* celldofs, i64 ncells, int nd, i64 ndofs, DofGather* out) {
  const i64 n = ncells * nd;
  out->ndofs = ndofs; out->ncontrib = n;
  GRMP_TRY(out->segptr.alloc(ndofs + 1));
  if (n >= 4294967295ll) return fail(GRMP_EUNSUPPORTED, "more than 2^32 local vector entries");
  if (n == 0) {
    segptr_from_sorted_u32<<<nblk(ndofs + 1), 256, 0, s>>>(nullptr, 0, ndofs, out->segptr.p);
    GRMP_CUDA(cudaGetLastError());
    return GRMP_OK;
  }
  GRMP_TRY(out->gsrc.alloc(n)); GRMP_TRY(out->gcell.alloc(n));
  DevBuf<u32> k1, k2, v1, v2; DevBuf<unsigned char> temp;
  GRMP_TRY(k1.alloc(n)); GRMP_TRY(k2.alloc(n)); GRMP_TRY(v1.alloc(n)); GRMP_TRY(v2.alloc(n));
  dof_keys<<<nblk(n), 256, 0, s>>>(celldofs, n, k1.p, v1.p);
  cub::DoubleBuffer<u32> dk(k1.p, k2.p), dv(v1.p, v2.p);
  size_t tb = 0;
  const int end_bit = bits_for((u64)ndofs) + 1;
  GRMP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, n, 0, end_bit > 32 ? 32 : end_bit, s));
  GRMP_TRY(temp.alloc(tb));
  GRMP_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tb, dk, dv, n, 0, end_bit > 32 ? 32 : end_bit, s));
  dof_fill<<<nblk(n), 256, 0, s>>>(dv.Current(), n, ncells, nd, out->gsrc.p, out->gcell.p);
  segptr_from_sorted_u32<<<nblk(ndofs + 1), 256, 0, s>>>(dk.Current(), n, ndofs, out->segptr.p);
  GRMP_CUDA(cudaGetLastError());
  GRMP_CUDA(cudaStreamSynchronize(s));
  return GRMP_OK;
}

int launch_lf_gather(cudaStream_t s, const DofGather& dg, const double* lbuf, const unsigned char* active, double* b) {
  if (dg.ndofs == 0) return GRMP_OK;
  lf_gather_kernel<<<nblk(dg.ndofs), 256, 0, s>>>(dg.segptr.p, dg.gsrc.p, dg.gcell.p, lbuf, active, dg.ndofs, b);
  GRMP_CUDA(cudaGetLastError());
  return GRMP_OK;
}

}  // namespace grmp
