// colpath.cu -- owner-computes column kernels: the numeric kernel family for every (element, operator) pair on the path.
//
// Replaces the cell loop of assemble! (bilinearform.jl:226-377) on a frozen pattern.  One THREAD owns one matrix column
// (a dof of the column space); it walks the cells that contain the dof, evaluates its own local column of every such cell
// (colpath_ev.cuh) and accumulates the entries in the shared-memory image of its column.  The images of the 32 columns of a
// warp are interleaved ([slot][lane]), so the read-modify-write of a warp never has a bank conflict whatever the slots are;
// finally every lane streams its column to nzval with 32-byte stores (full sectors).  Every stored non-zero is written
// exactly once, nothing is read-modify-written in global memory, there are no atomics and the summation order is fixed
// (cells ascending, like the reference) -> deterministic.
// Columns are processed in a locality order (sorted by their first cell, then by length inside a tile), not in dof order:
// dofs that are numbered far apart but live on the same cells (vertices of a refined grid) share the geometry cache.
//
// Data a CTA (tile = NW groups of 32 columns) touches:
//   * per-cell geometry records (update_trafo!/mapderiv!, feevaluator.jl:371-390, plus face signs / normals), evaluated ONCE
//     per cell and assembly by cell_geo_kernel into a global array that the column threads read through L1/L2 (every cell is
//     read by its nd columns; the locality order keeps those reads close in time).  Up to round 2 the geometry was rebuilt per
//     tile into a shared-memory cache: a five-deep chain of dependent global loads and a CTA barrier in front of ~3 rounds of
//     work per thread -- a third of the kernel's time (profiles/r2_col_kernels_ncu.md);
//   * pair records (16/32/48 B): cell (24 bit), local function of the column, and for every local row its slot inside
//     the column (255: the entry is not in the pattern -- _addnz skipped it, fematrix.jl:54-58).  Records of a group are
//     stored ROUND-major (pair k of all 32 columns, then pair k+1, ...) so every round is one coalesced load;
//   * the column function's reference table in shared memory (lane-varying index), the row table in constant memory
//     (uniform index), quadrature weights in constant memory.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "colpath.cuh"
#include "colpath_ev.cuh"

namespace grmp {

namespace {

__constant__ double c_tabR[TABR_MAX];   // row table [q][s][a]
__constant__ double c_wq[WQ_MAX];

struct ColParams {
  GridView g;
  const uint4* recs;
  const unsigned short* pos_np;    // per position (column in locality order): pairs (cells)
  const unsigned char* pos_len;    //   stored entries
  const i64* pos_start;            //   first nzval slot of the column (0-based)
  const i64* pos_recbeg;           //   first record (exclusive scan of pos_np); groups start at multiples of 32
  const double* geo;        // geometry records of this assembly (cell_geo_kernel): [ncells][quads] then [ncells][tail pair], see load_record
  const u32* tile_list;     // tiles of this launch (one class)
  const double* tabC;
  double factor;
  double act_p[2];
  double* nzval;
  i64 ncols_used, ngroups;
  int nw, nq;
};

template <int NV> __device__ __forceinline__ u32 rec_byte(const u32 (&w)[NV * 4], int i) { return (w[i >> 2] >> (8 * (i & 3))) & 255u; }
// Pair record: bytes 0-3 cell | local function << 24, bytes 4 .. 4 + NROW slot of every local row inside the column (255: not in
// the pattern) and, where the record has room, NROW bits "this is the FIRST contribution to the slot": the first contribution is
// a plain store into the column image (no load, no zero-initialisation pass) -- a fifth of the shared-memory wavefronts of these
// LSU-bound kernels.
__host__ __device__ constexpr bool rec_has_first(int nrow, int nv) { return 4 + nrow + (nrow + 7) / 8 <= 16 * nv; }

// geometry record of every cell: [0] CellVolumes, then the row / column evaluator's data (CacheLayout)
template <class RowEv, class ColEv>
__global__ void __launch_bounds__(256) cell_geo_kernel(const GridView g, double* __restrict__ geo) {
  using L = CacheLayout<RowEv, ColEv>;
  const i64 cell = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (cell >= g.ncells) return;
  double cr[L::STRIDE];
#pragma unroll
  for (int i = 0; i < L::STRIDE; i++) cr[i] = 0.0;
  build_cell_cache<RowEv, ColEv>(g, cell, 1.0, cr);
  store_record<L::STRIDE>(geo, g.ncells, cell, cr);
}

template <class RowEv, class ColEv, int ACT, int NV, int NQ>
__global__ void __launch_bounds__(256) col_kernel(const ColParams p) {
  using L = CacheLayout<RowEv, ColEv>;
  static_assert(RowEv::RD == ColEv::RD, "operator result dimensions must match");
  static_assert(RowEv::ED == ColEv::ED, "one grid");
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthr = blockDim.x;
  constexpr int nq = NQ;
  constexpr int ntab = ColEv::NAS * nq * CT_PAD;
  __shared__ u32 s_maxlen[8];
  double* const sCt = sm;
  double* const img = sm + ((ntab + 1) & ~1);
  const i64 tile = p.tile_list[blockIdx.x];
  const i64 grp = tile * p.nw + warp;
  const i64 pos = grp * 32 + lane;
  const bool has = grp < p.ngroups && pos < p.ncols_used;
  const u32 np = has ? p.pos_np[pos] : 0u;
  const u32 len = has ? p.pos_len[pos] : 0u;
  const i64 recbase = grp < p.ngroups ? p.pos_recbeg[grp * 32] : 0;
  double* const dst = p.nzval + (has ? p.pos_start[pos] : 0);
  const u32 maxlen = __reduce_max_sync(0xffffffffu, len);
  if (lane == 0) s_maxlen[warp] = maxlen;
  for (int i = tid; i < ntab; i += nthr) sCt[i] = p.tabC[i];
  const u32 maxnp = __reduce_max_sync(0xffffffffu, np);
  const u32 lt = (1u << lane) - 1u;
  u32 rbase = 0;
  auto next_idx = [&](u32 k) -> u32 {      // record of (lane, round k); must be called for k = 0, 1, 2, ... by the whole warp
    const u32 bal = __ballot_sync(0xffffffffu, k < np);
    const u32 idx = rbase + __popc(bal & lt);
    rbase += __popc(bal);
    return idx;
  };
  // record and geometry of round 0 are in flight while the table is staged
  uint4 rn[NV];                            // record of the next round (prefetched one round ahead)
  {
    const u32 i0 = next_idx(0);
#pragma unroll
    for (int v = 0; v < NV; v++) rn[v] = make_uint4(0xff000000u, 0, 0, 0);
    if (0 < np) {
#pragma unroll
      for (int v = 0; v < NV; v++) rn[v] = __ldg(p.recs + (size_t)(recbase + i0) * NV + v);
    }
  }
  __syncthreads();
  if (grp >= p.ngroups) return;
  u32 acc_off = 0;
  for (int w2 = 0; w2 < warp; w2++) acc_off += s_maxlen[w2] + 1u;     // + 1: trash row of every warp (entries that are not in the pattern)
  // image of the warp's 32 columns: slot k of lane l at [k][l] -> every warp access touches 32 consecutive doubles
  double* const a = img + (size_t)acc_off * 32 + lane;
  constexpr bool FT = rec_has_first(RowEv::NROW, NV);
  if (!FT)
    for (u32 k = 0; k < maxlen; k++) a[k * 32] = 0.0;
  for (u32 k = 0; k < maxnp; k++) {
    u32 w[NV * 4];
#pragma unroll
    for (int v = 0; v < NV; v++) { w[4 * v] = rn[v].x; w[4 * v + 1] = rn[v].y; w[4 * v + 2] = rn[v].z; w[4 * v + 3] = rn[v].w; }
    const u32 lc = w[0] >> 24;
    const bool work = k < np && lc != 255u;
    // geometry record of this round's cell: 256-bit loads, issued before the next record is requested
    double cr[L::STRIDE];
    if (work) {
      load_record<L::STRIDE>(p.geo, p.g.ncells, (i64)(w[0] & 0xffffffu), cr);
    }
    {
      const u32 i1 = next_idx(k + 1);
      if (k + 1 < np) {
#pragma unroll
        for (int v = 0; v < NV; v++) rn[v] = __ldg(p.recs + (size_t)(recbase + i1) * NV + v);
      }
    }
    if (work) {
      const double s = cr[0] * p.factor;
      typename RowEv::Regs RR;
      typename ColEv::Regs RC;
      RowEv::load(cr + L::OFF_R, RR);
      ColEv::load(cr + L::OFF_C, RC);
      typename RowEv::Acc A;
      RowEv::acc_zero(A);
      typename ColEv::Prep PC;
      ColEv::prep(RC, (int)lc, PC);
#pragma unroll 1
      for (int q = 0; q < nq; q++) {
        double Y[ColEv::RD];
        ColEv::col_eval(RC, PC, sCt, nq, q, Y);
        const double ws = c_wq[q] * s;
#pragma unroll
        for (int i = 0; i < ColEv::RD; i++) Y[i] *= ws;
        apply_action_col<ACT, ColEv::RD>(p.act_p, Y);
        double U[RowEv::NCU][RowEv::NAS];
        RowEv::pullback(RR, Y, U);
        RowEv::acc_rows(A, U, c_tabR + q * (RowEv::NSF * RowEv::NAS));
      }
      // rows that are not in the pattern (slot 255) go to the warp's trash row: no branch per row
      RowEv::emit_rows(RR, A, [&](int r, double v) {
        const u32 o = min(rec_byte<NV>(w, 4 + r), maxlen);
        if (FT) {
          const bool first = (rec_byte<NV>(w, 4 + RowEv::NROW + (r >> 3)) >> (r & 7)) & 1u;
          const double old = first ? 0.0 : a[o * 32];
          a[o * 32] = old + v;
        } else a[o * 32] += v;
      });
    }
  }
  // write-out: every lane streams its own column; 32-byte stores cover whole sectors, the unaligned ends go element-wise
  if (maxlen == 0) return;
  u32 head = (4u - (u32)(((size_t)dst >> 3) & 3u)) & 3u;
  if (head > len) head = len;
#pragma unroll
  for (u32 i = 0; i < 3; i++)
    if (i < head) dst[i] = a[i * 32];
  const u32 nbody = (len - head) >> 2;
  const u32 maxbody = __reduce_max_sync(0xffffffffu, nbody);
  for (u32 b = 0; b < maxbody; b++) {
    if (b < nbody) {
      const u32 k = head + 4 * b;
      const double v0 = a[k * 32], v1 = a[(k + 1) * 32], v2 = a[(k + 2) * 32], v3 = a[(k + 3) * 32];
      asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + k), "d"(v0), "d"(v1), "d"(v2), "d"(v3) : "memory");
    }
  }
  const u32 kt = head + 4 * nbody;
#pragma unroll
  for (u32 i = 0; i < 3; i++)
    if (kt + i < len) dst[kt + i] = a[(kt + i) * 32];
}

// ---- closed-form kernels (colpath_ev.cuh, "reference tensor" evaluators): same tiles, records, image and write-out as
//      col_kernel; the quadrature loop is replaced by NT multiply-adds per local row against K_t[s][s_col] in shared memory --------
template <class F>
__global__ void __launch_bounds__(256) cf_geo_kernel(const GridView g, double act0, double act1, double* __restrict__ geo) {
  const i64 cell = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (cell >= g.ncells) return;
  double cr[F::STRIDE];
#pragma unroll
  for (int i = 0; i < F::STRIDE; i++) cr[i] = 0.0;
  const double act_p[2] = {act0, act1};
  F::geo(g, cell, act_p, cr);
  store_record<F::STRIDE>(geo, g.ncells, cell, cr);
}

template <class F, int NV>
__global__ void __launch_bounds__(256) cf_kernel(const ColParams p) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthr = blockDim.x;
  constexpr int ntab = F::NSF * cf_kblock(F::NSF, F::NT);
  __shared__ u32 s_maxlen[8];
  double* const sK = sm;
  double* const img = sm + ntab;
  const i64 tile = p.tile_list[blockIdx.x];
  const i64 grp = tile * p.nw + warp;
  const i64 pos = grp * 32 + lane;
  const bool has = grp < p.ngroups && pos < p.ncols_used;
  const u32 np = has ? p.pos_np[pos] : 0u;
  const u32 len = has ? p.pos_len[pos] : 0u;
  const i64 recbase = grp < p.ngroups ? p.pos_recbeg[grp * 32] : 0;
  double* const dst = p.nzval + (has ? p.pos_start[pos] : 0);
  const u32 maxlen = __reduce_max_sync(0xffffffffu, len);
  if (lane == 0) s_maxlen[warp] = maxlen;
  for (int i = tid; i < ntab; i += nthr) sK[i] = p.tabC[i];
  const u32 maxnp = __reduce_max_sync(0xffffffffu, np);
  const u32 lt = (1u << lane) - 1u;
  u32 rbase = 0;
  auto next_idx = [&](u32 k) -> u32 {
    const u32 bal = __ballot_sync(0xffffffffu, k < np);
    const u32 idx = rbase + __popc(bal & lt);
    rbase += __popc(bal);
    return idx;
  };
  uint4 rn[NV];
  {
    const u32 i0 = next_idx(0);
#pragma unroll
    for (int v = 0; v < NV; v++) rn[v] = make_uint4(0xff000000u, 0, 0, 0);
    if (0 < np) {
#pragma unroll
      for (int v = 0; v < NV; v++) rn[v] = __ldg(p.recs + (size_t)(recbase + i0) * NV + v);
    }
  }
  __syncthreads();
  if (grp >= p.ngroups) return;
  u32 acc_off = 0;
  for (int w2 = 0; w2 < warp; w2++) acc_off += s_maxlen[w2] + 1u;
  double* const a = img + (size_t)acc_off * 32 + lane;
  constexpr bool FT = rec_has_first(F::NROW, NV);
  if (!FT)
    for (u32 k = 0; k < maxlen; k++) a[k * 32] = 0.0;
  const double factor = p.factor;
  for (u32 k = 0; k < maxnp; k++) {
    u32 w[NV * 4];
#pragma unroll
    for (int v = 0; v < NV; v++) { w[4 * v] = rn[v].x; w[4 * v + 1] = rn[v].y; w[4 * v + 2] = rn[v].z; w[4 * v + 3] = rn[v].w; }
    const u32 lc = w[0] >> 24;
    const bool work = k < np && lc != 255u;
    double cr[F::STRIDE];
    if (work) {
      load_record<F::STRIDE>(p.geo, p.g.ncells, (i64)(w[0] & 0xffffffu), cr);
    }
    {
      const u32 i1 = next_idx(k + 1);
      if (k + 1 < np) {
#pragma unroll
        for (int v = 0; v < NV; v++) rn[v] = __ldg(p.recs + (size_t)(recbase + i1) * NV + v);
      }
    }
    if (work) {
      F::column(cr, (int)lc, sK, [&](int r, double v) {
        const u32 o = min(rec_byte<NV>(w, 4 + r), maxlen);     // rows that are not in the pattern go to the warp's trash row
        if (FT) {
          const bool first = (rec_byte<NV>(w, 4 + F::NROW + (r >> 3)) >> (r & 7)) & 1u;
          const double old = first ? 0.0 : a[o * 32];
          a[o * 32] = fma(v, factor, old);
        } else a[o * 32] = fma(v, factor, a[o * 32]);
      });
    }
  }
  if (maxlen == 0) return;
  u32 head = (4u - (u32)(((size_t)dst >> 3) & 3u)) & 3u;
  if (head > len) head = len;
#pragma unroll
  for (u32 i = 0; i < 3; i++)
    if (i < head) dst[i] = a[i * 32];
  const u32 nbody = (len - head) >> 2;
  const u32 maxbody = __reduce_max_sync(0xffffffffu, nbody);
  for (u32 b = 0; b < maxbody; b++) {
    if (b < nbody) {
      const u32 k = head + 4 * b;
      const double v0 = a[k * 32], v1 = a[(k + 1) * 32], v2 = a[(k + 2) * 32], v3 = a[(k + 3) * 32];
      asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst + k), "d"(v0), "d"(v1), "d"(v2), "d"(v3) : "memory");
    }
  }
  const u32 kt = head + 4 * nbody;
#pragma unroll
  for (u32 i = 0; i < 3; i++)
    if (kt + i < len) dst[kt + i] = a[(kt + i) * 32];
}

// ---- one-time record build ------------------------------------------------------------------------------------------------
struct PackParams {
  const i64* colptr; const i64* rowval;        // pattern, 1-based
  const i64* pairbeg; const u32* gcell; const u32* gsrc;   // column -> (cell, local dof) pairs, cells ascending
  const u32* colperm;                          // position -> column
  const i64* pos_recbeg;                       // position -> first record
  const i32* dofsR; int ndR;                   // CellDofs of the row space
  const i32* orient; const i32* regions; RegionFilter reg;
  i64 ncells, ncols_used;
  int nrow;            // rows of the kernel's row loop (reference functions for BDM1 3D)
  int row_bdm3, col_bdm3;   // BDM1 3D: local dof <-> reference function through CellFaceOrientations (hdiv_bdm1.jl:313-328)
  int nw, nv;
  uint4* recs;
  int* err;
};

__device__ __forceinline__ int bdm3_ref_of_local(const i32* o, int l) {   // subset[l], 0-based
  const int j = l / 3, m = l - 3 * j, oj = o[j] - 1;
  const int s1 = oj == 0 ? 1 : (oj == 1 ? 0 : (oj == 2 ? 1 : 2)), s2 = oj == 0 ? 2 : (oj == 1 ? 2 : (oj == 2 ? 0 : 1));
  return m == 0 ? 4 * j : (m == 1 ? 4 * j + 3 - s1 : 4 * j + 3 - s2);
}
__device__ __forceinline__ int bdm3_local_of_ref(const i32* o, int r) {   // -1: not selected on this cell
  const int j = r >> 2;
  for (int m = 0; m < 3; m++)
    if (bdm3_ref_of_local(o, 3 * j + m) == r) return 3 * j + m;
  return -1;
}

// sort key of the locality order: first (lowest) cell of the column, then the column itself
__global__ void order_keys(const i64* colptr, const i64* pairbeg, const u32* gcell, i64 ncols_used, u64* keys, u32* ids, int* err) {
  const i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (j >= ncols_used) return;
  const i64 l = colptr ? colptr[j + 1] - colptr[j] : 0, n = pairbeg[j + 1] - pairbeg[j];
  if (l > 254 || n > 65535) atomicExch(err, 1);
  const u64 first = n > 0 ? (u64)gcell[pairbeg[j]] : 0xffffffffull;
  keys[j] = (first << 32) | (u64)j;
  ids[j] = (u32)j;
}
// second key: (tile, length) -> columns of similar length share a warp (the interleaved image is sized by the longest)
__global__ void tile_len_keys(const u32* order1, const i64* colptr, i64 ncols_used, int cpt, u64* keys) {
  const i64 pos = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (pos >= ncols_used) return;
  const i64 j = order1[pos];
  keys[pos] = ((u64)(pos / cpt) << 8) | (u64)(colptr[j + 1] - colptr[j]);
}
__global__ void pos_arrays(const u32* colperm, const i64* colptr, const i64* pairbeg, i64 ncols_used, unsigned short* pos_np, unsigned char* pos_len,
                           i64* pos_start, i64* np64) {
  const i64 pos = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (pos >= ncols_used) return;
  const i64 j = colperm[pos];
  const i64 n = pairbeg[j + 1] - pairbeg[j];
  pos_np[pos] = (unsigned short)n; pos_len[pos] = (unsigned char)(colptr[j + 1] - colptr[j]); pos_start[pos] = colptr[j] - 1;
  np64[pos] = n;
}
__global__ void tile_keys(const u32* colperm, const i64* pairbeg, const i64* pos_recbeg, const u32* gcell, i64 ncols_used, int cpt, u64* keys) {
  const i64 pos = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (pos >= ncols_used) return;
  const i64 j = colperm[pos];
  const u64 t = (u64)(pos / cpt) << 32;
  const i64 kb = pairbeg[j], n = pairbeg[j + 1] - kb, ob = pos_recbeg[pos];
  for (i64 k = 0; k < n; k++) keys[ob + k] = t | gcell[kb + k];
}
__global__ void tile_ptr(const u64* uniq, i64 n, i64 ntiles, u32* tile_cellptr, int* max_cells) {
  const i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t > ntiles) return;
  auto lb = [&](u64 key) {
    i64 lo = 0, hi = n;
    while (lo < hi) { const i64 mid = (lo + hi) >> 1; if (uniq[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
  };
  const i64 b = lb((u64)t << 32);
  tile_cellptr[t] = (u32)b;
  if (t < ntiles) atomicMax(max_cells, (int)(lb((u64)(t + 1) << 32) - b));
}
// shared-memory doubles a tile needs: column table + the interleaved images of its groups
__global__ void tile_need(const unsigned char* pos_len, i64 ntiles, i64 ncols_used, int nw, int ntab_even, int* need) {
  const i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  i64 n = ntab_even;
  for (int w = 0; w < nw; w++) {
    const i64 p0 = min((t * nw + w) * 32, ncols_used), p1 = min(p0 + 32, ncols_used);
    int mx = 0;
    for (i64 q = p0; q < p1; q++) mx = max(mx, (int)pos_len[q]);
    n += 32 * (i64)(mx + 1);
  }
  need[t] = (int)min(n, (i64)0x7fffffff);
}
__global__ void low32(const u64* a, i64 n, u32* out) {
  const i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (u32)(a[i] & 0xffffffffull);
}

__global__ void pack_records(PackParams p) {
  const i64 grp = (blockIdx.x * (i64)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (grp * 32 >= p.ncols_used) return;
  const i64 pos = grp * 32 + lane;
  const bool has = pos < p.ncols_used;
  const i64 j = has ? (i64)p.colperm[pos] : 0;
  const i64 kb = has ? p.pairbeg[j] : 0;
  const u32 np = has ? (u32)(p.pairbeg[j + 1] - kb) : 0u;
  const i64 recbase = p.pos_recbeg[grp * 32];
  const i64 cb = has ? p.colptr[j] - 1 : 0, ce = has ? p.colptr[j + 1] - 1 : 0;
  const u32 maxnp = __reduce_max_sync(0xffffffffu, np);
  const u32 lt = (1u << lane) - 1u;
  u32 rbase = 0;
  const bool ft = rec_has_first(p.nrow, p.nv);
  u32 touched[8] = {0, 0, 0, 0, 0, 0, 0, 0};          // slots of the column that an earlier pair has written
  for (u32 k = 0; k < maxnp; k++) {
    const u32 bal = __ballot_sync(0xffffffffu, k < np);
    const u32 idx = rbase + __popc(bal & lt);
    rbase += __popc(bal);
    if (k >= np) continue;
    const i64 cell = p.gcell[kb + k];
    int lc = (int)(p.gsrc[kb + k] / (u32)p.ncells);
    bool active = true;
    if (p.reg.n > 0) {
      active = false;
      if (p.regions) for (int r = 0; r < p.reg.n; r++) active = active || p.regions[cell] == p.reg.r[r];
    }
    if (p.col_bdm3) lc = bdm3_ref_of_local(p.orient + cell * 4, lc);
    u32 w[12];
    for (int i = 0; i < 12; i++) w[i] = 0xffffffffu;
    w[0] = (u32)cell | ((active ? (u32)lc : 255u) << 24);
    for (int r = 0; r < p.nrow; r++) {
      int l = r;
      if (p.row_bdm3) l = bdm3_local_of_ref(p.orient + cell * 4, r);
      u32 off = 255u;
      if (l >= 0) {
        const i64 target = p.dofsR[cell * p.ndR + l];     // 1-based row
        i64 a = cb, b = ce;
        while (a < b) { const i64 mid = (a + b) >> 1; if (p.rowval[mid] < target) a = mid + 1; else b = mid; }
        if (a < ce && p.rowval[a] == target) off = (u32)(a - cb);
      }
      const int byte = 4 + r;
      w[byte >> 2] = (w[byte >> 2] & ~(255u << (8 * (byte & 3)))) | (off << (8 * (byte & 3)));
      if (ft) {      // all-ones default = "first" (rows outside the pattern go to the trash row: a store is as good as an add)
        const bool first = off == 255u || !active || !((touched[off >> 5] >> (off & 31)) & 1u);
        if (off != 255u && active) touched[off >> 5] |= 1u << (off & 31);
        const int fb = 4 + p.nrow + (r >> 3);
        if (!first) w[fb >> 2] &= ~(1u << (8 * (fb & 3) + (r & 7)));
      }
    }
    uint4* dst = p.recs + (size_t)(recbase + idx) * p.nv;
    for (int v = 0; v < p.nv; v++) dst[v] = make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
  }
}

inline unsigned nblk(i64 n, int t = 256) { return (unsigned)((n + t - 1) / t); }

// ---- kernel table ----------------------------------------------------------------------------------------------------------
typedef int (*LaunchFn)(const ColParams&, int nblocks, int nthreads, int smem, cudaStream_t);
typedef int (*GeoFn)(const GridView&, double* geo, cudaStream_t);
struct Variant {
  bool (*match)(const ColEvalDesc& row, const ColEvalDesc& col, int act, int nq);
  LaunchFn launch;
  GeoFn geo;
  int (*cache_stride)();
  int nrow, nv, tabR_per_q, nas_c;
};

template <class RowEv, class ColEv, int ACT, int NQ> struct VariantImpl {
  static constexpr int NV = (4 + RowEv::NROW + 15) / 16;
  static bool match(const ColEvalDesc& row, const ColEvalDesc& col, int act, int nq) { return act == ACT && nq == NQ && ev_matches<RowEv>(row) && ev_matches<ColEv>(col); }
  static int launch(const ColParams& p, int nblocks, int nthreads, int smem, cudaStream_t s) {
    static int attr_set = 0;
    if (smem > attr_set) {
      GRMP_CUDA(cudaFuncSetAttribute(col_kernel<RowEv, ColEv, ACT, NV, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_set = smem;
    }
    col_kernel<RowEv, ColEv, ACT, NV, NQ><<<nblocks, nthreads, smem, s>>>(p);
    GRMP_CUDA(cudaGetLastError());
    return GRMP_OK;
  }
  static int geo(const GridView& g, double* out, cudaStream_t s) {
    cell_geo_kernel<RowEv, ColEv><<<(unsigned)((g.ncells + 255) / 256), 256, 0, s>>>(g, out);
    GRMP_CUDA(cudaGetLastError());
    return GRMP_OK;
  }
  static int cache_stride() { return CacheLayout<RowEv, ColEv>::STRIDE; }
};
#define GRMP_UNPAREN(...) __VA_ARGS__
#define GRMP_VARIANT(R, C, A, Q)                                                                                       \
  {&VariantImpl<GRMP_UNPAREN R, GRMP_UNPAREN C, A, Q>::match, &VariantImpl<GRMP_UNPAREN R, GRMP_UNPAREN C, A, Q>::launch, \
   &VariantImpl<GRMP_UNPAREN R, GRMP_UNPAREN C, A, Q>::geo, &VariantImpl<GRMP_UNPAREN R, GRMP_UNPAREN C, A, Q>::cache_stride, GRMP_UNPAREN R::NROW, VariantImpl<GRMP_UNPAREN R, GRMP_UNPAREN C, A, Q>::NV, \
   GRMP_UNPAREN R::NSF * GRMP_UNPAREN R::NAS, GRMP_UNPAREN C::NAS},
const Variant VARIANTS[] = {GRMP_SQUARE_FORMS(GRMP_VARIANT) GRMP_RECT_FORMS(GRMP_VARIANT)};
constexpr int NVARIANTS = sizeof(VARIANTS) / sizeof(VARIANTS[0]);

typedef int (*CfGeoFn)(const GridView&, const double* act_p, double* geo, cudaStream_t);
struct CfVariant {
  bool (*match)(const ColEvalDesc& row, const ColEvalDesc& col, int act);
  LaunchFn launch;
  CfGeoFn geo;
  int stride, nrow, nv, nt, nsf, ed, symk;
};
template <class F> struct CfVariantImpl {
  static constexpr int NV = (4 + F::NROW + 15) / 16;
  static bool match(const ColEvalDesc& row, const ColEvalDesc& col, int act) {
    return act == F::ACT && ev_matches<typename F::Quad>(row) && ev_matches<typename F::Quad>(col);
  }
  static int launch(const ColParams& p, int nblocks, int nthreads, int smem, cudaStream_t s) {
    static int attr_set = 0;
    if (smem > attr_set) {
      GRMP_CUDA(cudaFuncSetAttribute(cf_kernel<F, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_set = smem;
    }
    cf_kernel<F, NV><<<nblocks, nthreads, smem, s>>>(p);
    GRMP_CUDA(cudaGetLastError());
    return GRMP_OK;
  }
  static int geo(const GridView& g, const double* act_p, double* out, cudaStream_t s) {
    cf_geo_kernel<F><<<(unsigned)((g.ncells + 255) / 256), 256, 0, s>>>(g, act_p[0], act_p[1], out);
    GRMP_CUDA(cudaGetLastError());
    return GRMP_OK;
  }
};
#define GRMP_CFV(F) {&CfVariantImpl<GRMP_UNPAREN F>::match, &CfVariantImpl<GRMP_UNPAREN F>::launch, &CfVariantImpl<GRMP_UNPAREN F>::geo, \
                     GRMP_UNPAREN F::STRIDE, GRMP_UNPAREN F::NROW, CfVariantImpl<GRMP_UNPAREN F>::NV, GRMP_UNPAREN F::NT, GRMP_UNPAREN F::NSF, \
                     GRMP_UNPAREN F::ED, GRMP_UNPAREN F::SYMK ? 1 : 0},
const CfVariant CFVARIANTS[] = {GRMP_CF_FORMS(GRMP_CFV)};
constexpr int NCFVARIANTS = sizeof(CFVARIANTS) / sizeof(CFVARIANTS[0]);

int describe(const EvalView& e, int ed, ColEvalDesc* d) {
  d->op = e.op; d->ed = ed; d->nd = e.nd; d->fam = e.fam; d->nbub = 0;
  if (e.op == GRMP_OP_RECON_ID_RT0 || e.op == GRMP_OP_RECON_ID_BDM1) {
    if (ed != 2 || e.fam != FAM_H1BR) return -1;       // tetrahedra: quadrature order 6 (46 points) stays on the generic path
    d->kind = 2; d->nc = 1; d->nds = e.tab_nd;
    return 0;
  }
  if (e.fam == FAM_H1) {
    if (e.nd % e.ncomp) return -1;
    d->kind = 0; d->nc = e.ncomp; d->nds = e.nd / e.ncomp;
  } else if (e.fam == FAM_H1BR) {
    d->kind = 0; d->nc = ed; d->nds = ed + 1; d->nbub = ed + 1;
  } else if (e.fam == FAM_RT0 || e.fam == FAM_BDM1) {
    d->kind = 1; d->nc = 1; d->nds = e.nd_all;
  } else return -1;
  return 0;
}

// host: scalar row / column tables from the caller's evaluator tables; verifies the componentwise structure
int make_tables(const ColEvalDesc& d, int nq, const std::vector<double>& vals, const std::vector<double>& derivs, int nd_all, int ncomp,
                std::vector<double>* T /* [s][a][q] */, int* nas) {
  const int ed = d.ed;
  const bool der = (d.op != GRMP_OP_ID && d.kind != 2);
  const bool values_table = (d.op == GRMP_OP_ID || d.kind == 2);          // Hdiv Identity / reconstruction: value table of the Hdiv space
  *nas = (d.kind == 0) ? (der ? ed : 1) : (values_table ? ed : 1);
  const int nsf = d.nds + d.nbub;
  T->assign((size_t)nsf * (*nas) * nq, 0.0);
  auto V = [&](int q, int l, int c) { return vals[((size_t)q * nd_all + l) * ncomp + c]; };
  auto D = [&](int q, int a, int l, int c) { return derivs[((size_t)q * ed + a) * ((size_t)nd_all * ncomp) + l + (size_t)c * nd_all]; };
  if (der && derivs.size() != (size_t)nq * ed * nd_all * ncomp) return fail(GRMP_EUNSUPPORTED, "column kernels: derivative table missing");
  if (!der && vals.size() != (size_t)nq * nd_all * ncomp) return fail(GRMP_EUNSUPPORTED, "column kernels: value table missing");
  if (d.kind != 0) {
    for (int r = 0; r < nsf; r++) for (int q = 0; q < nq; q++) {
      if (values_table) for (int a = 0; a < ed; a++) (*T)[((size_t)r * ed + a) * nq + q] = V(q, r, a);
      else { double s = 0.0; for (int jj = 0; jj < ed; jj++) s += D(q, jj, r, jj); (*T)[(size_t)r * nq + q] = s; }
    }
    return GRMP_OK;
  }
  const int ndof = d.nc * d.nds + d.nbub;
  if (ndof != nd_all || ncomp != d.nc) return fail(GRMP_EUNSUPPORTED, "column kernels: table shape");
  for (int s = 0; s < nsf; s++) {
    const int l0 = s < d.nds ? s : d.nc * d.nds + (s - d.nds);
    for (int q = 0; q < nq; q++) {
      if (der) for (int a = 0; a < ed; a++) (*T)[((size_t)s * ed + a) * nq + q] = D(q, a, l0, 0);
      else (*T)[(size_t)s * nq + q] = V(q, l0, 0);
    }
  }
  // structure check: dof (c, s) lives in component c only, bubbles carry the same scalar in every component
  for (int l = 0; l < ndof; l++) for (int c = 0; c < d.nc; c++) for (int q = 0; q < nq; q++) for (int a = 0; a < *nas; a++) {
    const bool bub = l >= d.nc * d.nds;
    const int s = bub ? d.nds + (l - d.nc * d.nds) : l % d.nds;
    const double expect = (bub || l / d.nds == c) ? (*T)[((size_t)s * (*nas) + a) * nq + q] : 0.0;
    const double got = der ? D(q, a, l, c) : V(q, l, c);
    if (got != expect) return fail(GRMP_EUNSUPPORTED, "column kernels: evaluator table is not componentwise");
  }
  return GRMP_OK;
}

}  // namespace

static u64 g_const_owner = 0;    // uid of the ColPath whose tables are in constant memory
static u64 g_next_uid = 1;

bool colpath_applicable(const BlfLocalParams& p, int nq, ColPath* cp) {
  if (p.apt == GRMP_APT_LUMPED) return false;
  const bool tr = (p.apt != GRMP_APT_SYMMETRIC && p.transposed);
  const EvalView& er = tr ? p.e2 : p.e1;
  const EvalView& ec = tr ? p.e1 : p.e2;
  if (describe(er, p.g.dim, &cp->row) || describe(ec, p.g.dim, &cp->col)) return false;
  if (nq > WQ_MAX) return false;
  cp->row_is_arg1 = !tr;
  cp->variant = -1;
  cp->cf_variant = -1;
  cp->nq = nq;
  // closed-form kernel where one exists (any quadrature rule: K is the rule's own sum; GRMP_COL_QUADRATURE=1 keeps the quadrature
  // kernels), and the quadrature kernel whose tables the cell-parallel alternatives share
  if (!getenv("GRMP_COL_QUADRATURE") && p.same_eval)
    for (int v = 0; v < NCFVARIANTS; v++)
      if (CFVARIANTS[v].match(cp->row, cp->col, p.action)) { cp->cf_variant = v; break; }
  for (int v = 0; v < NVARIANTS; v++)
    if (VARIANTS[v].match(cp->row, cp->col, p.action, nq) && (size_t)VARIANTS[v].tabR_per_q * nq <= TABR_MAX) { cp->variant = v; break; }
  // Reconstruction mass forms (R u, R v): the reconstructed functions are Piola images of P1 fields, the integrand has degree 2, but the
  // reference's rule follows the Bernardi-Raugel degree (order 4: 9 Stroud points).  Both rules are exact, so the quadrature sum equals
  // the one of ANY exact rule up to rounding: colpath_build factors the rule's own moment matrix into 3 virtual points (GRMP_COL_NO_REDUCE=1 disables)
  cp->variant_reduced = -1;
  if (cp->variant >= 0 && cp->row.kind == 2 && cp->col.kind == 2 && p.same_eval && nq > 3 && !getenv("GRMP_COL_NO_REDUCE"))
    for (int v = 0; v < NVARIANTS; v++)
      if (VARIANTS[v].match(cp->row, cp->col, p.action, 3)) { cp->variant_reduced = v; break; }
  return cp->variant >= 0 || cp->cf_variant >= 0;
}

int colpath_build(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, const std::vector<double>& w, const std::vector<double>& vals1,
                  const std::vector<double>& derivs1, const std::vector<double>& vals2, const std::vector<double>& derivs2, i64 ncols_owned,
                  bool quadrature_tables, ColPath* cp) {
  cudaStream_t s = ctx->stream;
  cp->built = false;
  cp->uid = g_next_uid++;
  if (quadrature_tables) { cp->cf_variant = -1; cp->variant_reduced = -1; }       // the cell-parallel kernels evaluate the caller's quadrature sum themselves
  if (cp->cf_variant < 0 && cp->variant < 0) return fail(GRMP_EUNSUPPORTED, "no column / cell kernel for this form and quadrature rule");
  const bool cf = cp->cf_variant >= 0;
  const bool tr = !cp->row_is_arg1;
  const EvalView& er = tr ? p.e2 : p.e1;
  const EvalView& ec = tr ? p.e1 : p.e2;
  int nq = p.nq;
  cp->nq = nq;
  // equivalent 3-point rule for reconstruction mass forms: K = sum_q w_q t_q t_q' (t_q = all table entries at point q) has rank <= 3
  // because every table row is affine in xref; three steps of a pivoted Cholesky factorisation give K = L L', i.e. three virtual
  // points with weight 1 whose "table values" are the columns of L.  Accepted only if the residual is at rounding level.
  std::vector<double> redT, redW;
  if (!cf && cp->variant_reduced >= 0) {
    std::vector<double> T;
    int nas = 0;
    GRMP_TRY(make_tables(cp->row, nq, vals1, derivs1, er.tab_nd, er.tab_nc, &T, &nas));
    const int nsf = cp->row.nds + cp->row.nbub, m = nsf * nas;
    std::vector<double> K((size_t)m * m), L((size_t)m * 3, 0.0);
    double kmax = 0.0;
    for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) {
      double v = 0.0;
      for (int q = 0; q < nq; q++) v += w[q] * T[(size_t)i * nq + q] * T[(size_t)j * nq + q];
      K[(size_t)i * m + j] = v;
      kmax = std::max(kmax, std::fabs(v));
    }
    std::vector<double> Rm = K;
    bool ok = kmax > 0.0;
    for (int k = 0; k < 3 && ok; k++) {
      int piv = 0;
      for (int i = 1; i < m; i++) if (Rm[(size_t)i * m + i] > Rm[(size_t)piv * m + piv]) piv = i;
      const double d = Rm[(size_t)piv * m + piv];
      if (!(d > 1e-10 * kmax)) { ok = (k > 0); break; }      // rank < 3 (e.g. constants only): fewer virtual points carry everything
      const double sd = std::sqrt(d);
      for (int i = 0; i < m; i++) L[(size_t)i * 3 + k] = Rm[(size_t)i * m + piv] / sd;
      for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) Rm[(size_t)i * m + j] -= L[(size_t)i * 3 + k] * L[(size_t)j * 3 + k];
    }
    double res = 0.0;
    for (size_t i = 0; i < Rm.size(); i++) res = std::max(res, std::fabs(Rm[i]));
    if (ok && res <= 1e-14 * kmax) {
      redT.assign((size_t)m * 3, 0.0);                       // [s][a][q'] like make_tables
      for (int i = 0; i < m; i++) for (int k = 0; k < 3; k++) redT[(size_t)i * 3 + k] = L[(size_t)i * 3 + k];
      redW.assign(3, 1.0);
      nq = 3;
      cp->nq = 3;
    } else {
      cp->variant_reduced = -1;                              // not an affine table set: keep the caller's rule
    }
  }
  const bool reduced = !cf && cp->variant_reduced >= 0;
  if (reduced) cp->variant = cp->variant_reduced;
  const Variant& V = VARIANTS[cf ? 0 : cp->variant];
  const i64 ncols = pat.ncols, ncells = p.g.ncells;
  const int v_nrow = cf ? CFVARIANTS[cp->cf_variant].nrow : V.nrow, v_nv = cf ? CFVARIANTS[cp->cf_variant].nv : V.nv;
  const int v_stride = cf ? CFVARIANTS[cp->cf_variant].stride : V.cache_stride();
  // tables
  if (cf) {
    // K_t[s][s_col] = sum_q w_q T[s][a][q] T[s_col][b][q] from the caller's own tables and weights (t = (a, b); symmetric G: a <= b and
    // K_ab + K_ba merged), layout [s_col][s][t] in blocks of cf_kblock doubles
    const CfVariant& F = CFVARIANTS[cp->cf_variant];
    std::vector<double> T;
    int nas = 0;
    GRMP_TRY(make_tables(cp->row, nq, vals1, derivs1, er.tab_nd, er.tab_nc, &T, &nas));
    const int nsf = F.nsf, ed = F.ed;
    if (nsf != cp->row.nds + cp->row.nbub) return fail(GRMP_EUNSUPPORTED, "column kernels: closed-form table shape");
    auto Kab = [&](int a, int b, int sI, int sJ) {
      double v = 0.0;
      for (int q = 0; q < nq; q++) v += w[q] * T[((size_t)sI * nas + a) * nq + q] * T[((size_t)sJ * nas + b) * nq + q];
      return v;
    };
    const int kb = cf_kblock(nsf, F.nt);
    std::vector<double> K((size_t)nsf * kb, 0.0);
    for (int t = 0; t < F.nt; t++) {
      int a = 0, b = 0;
      if (nas > 1) {
        if (F.symk) { a = sym_a(ed, t); b = sym_b(ed, t); } else { a = t / ed; b = t % ed; }
      }
      for (int sI = 0; sI < nsf; sI++) for (int sJ = 0; sJ < nsf; sJ++)
        K[(size_t)sJ * kb + sI * F.nt + t] = Kab(a, b, sI, sJ) + ((F.symk && a != b) ? Kab(b, a, sI, sJ) : 0.0);
    }
    if ((nas > 1) != (F.nt > 1)) return fail(GRMP_EUNSUPPORTED, "column kernels: closed-form table shape");
    GRMP_TRY(cp->tabC.upload(K.data(), K.size(), s));
    cp->tabR.clear();
    cp->wq = w;
  } else {
    const std::vector<double>& v2 = p.same_eval ? vals1 : vals2;
    const std::vector<double>& d2 = p.same_eval ? derivs1 : derivs2;
    std::vector<double> TR, TC;
    int nasR = 0, nasC = 0;
    if (reduced) {
      TR = redT; TC = redT;
      nasR = nasC = (int)(redT.size() / 3) / (cp->row.nds + cp->row.nbub);
    } else {
      GRMP_TRY(make_tables(cp->row, nq, tr ? v2 : vals1, tr ? d2 : derivs1, er.tab_nd, er.tab_nc, &TR, &nasR));
      GRMP_TRY(make_tables(cp->col, nq, tr ? vals1 : v2, tr ? derivs1 : d2, ec.tab_nd, ec.tab_nc, &TC, &nasC));
    }
    const int nsfR = cp->row.nds + cp->row.nbub, nsfC = cp->col.nds + cp->col.nbub;
    cp->tabR.assign((size_t)nq * nsfR * nasR, 0.0);        // [q][s][a]
    for (int q = 0; q < nq; q++) for (int sI = 0; sI < nsfR; sI++) for (int a = 0; a < nasR; a++)
      cp->tabR[((size_t)q * nsfR + sI) * nasR + a] = TR[((size_t)sI * nasR + a) * nq + q];
    std::vector<double> tc((size_t)nasC * nq * CT_PAD, 0.0);   // [a][q][CT_PAD]
    for (int a = 0; a < nasC; a++) for (int q = 0; q < nq; q++) for (int sI = 0; sI < nsfC; sI++)
      tc[((size_t)a * nq + q) * CT_PAD + sI] = TC[((size_t)sI * nasC + a) * nq + q];
    GRMP_TRY(cp->tabC.upload(tc.data(), tc.size(), s));
    cp->wq = reduced ? redW : w;
  }
  const i64 ncols_used = (ncols_owned >= 0 && ncols_owned < ncols) ? ncols_owned : ncols;
  cp->ncols_used = ncols_used;
  cp->ngroups = (ncols_used + 31) / 32;
  cp->nv = v_nv;
  if (ncols_used == 0 || pat.nnz == 0) { cp->ntiles = 0; cp->built = true; return GRMP_OK; }
  // (1) column -> (cell, local dof) pairs, cells ascending
  DofGather dg;
  GRMP_TRY(build_dofgather(s, ec.celldofs, ncells, ec.nd, ncols, &dg));
  i64 npairs_used = 0;
  GRMP_CUDA(cudaMemcpyAsync(&npairs_used, dg.segptr.p + ncols_used, 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  cp->npairs = npairs_used;
  // (2) locality order of the columns: by first cell (radix sort, stable)
  DevBuf<int> flags;                          // [0] error, [2] max tile cells
  GRMP_TRY(flags.alloc(4));
  GRMP_CUDA(cudaMemsetAsync(flags.p, 0, 16, s));
  DevBuf<u64> ck1, ck2; DevBuf<u32> ci1, ci2; DevBuf<unsigned char> temp;
  GRMP_TRY(ck1.alloc(ncols_used)); GRMP_TRY(ck2.alloc(ncols_used)); GRMP_TRY(ci1.alloc(ncols_used)); GRMP_TRY(ci2.alloc(ncols_used));
  order_keys<<<nblk(ncols_used), 256, 0, s>>>(pat.colptr.p, dg.segptr.p, dg.gcell.p, ncols_used, ck1.p, ci1.p, flags.p);
  GRMP_CUDA(cudaGetLastError());
  const bool natural = getenv("GRMP_COL_NATURAL_ORDER") != nullptr;    // experiment: keep the dof order
  {
    cub::DoubleBuffer<u64> dk(ck1.p, ck2.p);
    cub::DoubleBuffer<u32> dv(ci1.p, ci2.p);
    if (!natural) {
      size_t tb = 0;
      GRMP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, ncols_used, 0, 64, s));
      if (tb > temp.n) GRMP_TRY(temp.alloc(tb));
      GRMP_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tb, dk, dv, ncols_used, 0, 64, s));
    }
    if (dv.Current() != ci1.p) GRMP_CUDA(cudaMemcpyAsync(ci1.p, dv.Current(), (size_t)ncols_used * 4, cudaMemcpyDeviceToDevice, s));
  }
  // (3) tiles: NW groups per CTA; inside a tile columns are ordered by length; shrink the tile until the shared-memory image fits
  if (ncells > (i64)0x1000000) return fail(GRMP_EUNSUPPORTED, "column kernels: more than 2^24 cells on one device (24-bit cell ids in the records)");
  int nw = getenv("GRMP_COL_NW") ? atoi(getenv("GRMP_COL_NW")) : 4;
  nw = std::max(1, std::min(nw, 8));
  DevBuf<i64> np64;
  GRMP_TRY(cp->colperm.alloc(ncols_used)); GRMP_TRY(cp->pos_np.alloc(ncols_used)); GRMP_TRY(cp->pos_len.alloc(ncols_used));
  GRMP_TRY(cp->pos_start.alloc(ncols_used)); GRMP_TRY(cp->pos_recbeg.alloc(ncols_used + 1)); GRMP_TRY(np64.alloc(ncols_used + 1));
  int hflags[4] = {0, 0, 0, 0};
  for (;; nw >>= 1) {
    const int cpt = 32 * nw;
    const i64 ntiles = (ncols_used + cpt - 1) / cpt;
    {
      tile_len_keys<<<nblk(ncols_used), 256, 0, s>>>(ci1.p, pat.colptr.p, ncols_used, cpt, ck1.p);
      GRMP_CUDA(cudaGetLastError());
      GRMP_CUDA(cudaMemcpyAsync(ci2.p, ci1.p, (size_t)ncols_used * 4, cudaMemcpyDeviceToDevice, s));
      cub::DoubleBuffer<u64> dk(ck1.p, ck2.p);
      cub::DoubleBuffer<u32> dv(ci2.p, cp->colperm.p);
      size_t tb = 0;
      GRMP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, ncols_used, 0, 64, s));
      if (tb > temp.n) GRMP_TRY(temp.alloc(tb));
      GRMP_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tb, dk, dv, ncols_used, 0, 64, s));
      if (dv.Current() != cp->colperm.p) GRMP_CUDA(cudaMemcpyAsync(cp->colperm.p, dv.Current(), (size_t)ncols_used * 4, cudaMemcpyDeviceToDevice, s));
    }
    pos_arrays<<<nblk(ncols_used), 256, 0, s>>>(cp->colperm.p, pat.colptr.p, dg.segptr.p, ncols_used, cp->pos_np.p, cp->pos_len.p, cp->pos_start.p, np64.p);
    GRMP_CUDA(cudaGetLastError());
    GRMP_CUDA(cudaMemsetAsync(np64.p + ncols_used, 0, 8, s));
    {
      size_t tb = 0;
      GRMP_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, np64.p, cp->pos_recbeg.p, ncols_used + 1, s));
      if (tb > temp.n) GRMP_TRY(temp.alloc(tb));
      GRMP_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, tb, np64.p, cp->pos_recbeg.p, ncols_used + 1, s));
    }
    GRMP_CUDA(cudaMemcpyAsync(hflags, flags.p, 16, cudaMemcpyDeviceToHost, s));
    GRMP_CUDA(cudaStreamSynchronize(s));
    if (hflags[0]) return fail(GRMP_EUNSUPPORTED, "column kernels: a column has more than 254 entries or 65535 cells");
    // shared-memory need of every tile; tiles are launched in classes of similar need
    const int ntab_even = cf ? CFVARIANTS[cp->cf_variant].nsf * cf_kblock(CFVARIANTS[cp->cf_variant].nsf, CFVARIANTS[cp->cf_variant].nt) : (V.nas_c * nq * CT_PAD + 1) & ~1;
    DevBuf<int> need_d;
    GRMP_TRY(need_d.alloc(ntiles));
    tile_need<<<nblk(ntiles), 256, 0, s>>>(cp->pos_len.p, ntiles, ncols_used, nw, ntab_even, need_d.p);
    GRMP_CUDA(cudaGetLastError());
    std::vector<int> need(ntiles);
    GRMP_CUDA(cudaMemcpyAsync(need.data(), need_d.p, (size_t)ntiles * 4, cudaMemcpyDeviceToHost, s));
    GRMP_CUDA(cudaStreamSynchronize(s));
    int need_max = 0;
    for (int n : need) need_max = std::max(need_max, n);
    if ((i64)need_max * 8 <= 200 * 1024) {
      cp->nw = nw; cp->ntiles = ntiles; cp->smem_bytes = need_max * 8;
      // classes: capacities that let 16 / 8 / 4 / 2 / 1 CTAs share an SM (227 KB usable, 1 KB reserved per CTA)
      const int caps[6] = {6 * 1024, 13 * 1024, 27 * 1024, 55 * 1024, 112 * 1024, 200 * 1024};
      std::vector<std::vector<u32>> lists(6);
      for (i64 t = 0; t < ntiles; t++) {
        int c = 0;
        while (c < 5 && (i64)need[t] * 8 > caps[c]) c++;
        lists[c].push_back((u32)t);
      }
      std::vector<u32> all;
      cp->classes.clear();
      for (int c = 5; c >= 0; c--) {       // crowded tiles first
        if (lists[c].empty()) continue;
        int mx = 0;
        for (u32 t : lists[c]) mx = std::max(mx, need[t]);
        cp->classes.push_back(ColPath::TileClass{mx * 8, (i64)all.size(), (i64)lists[c].size()});
        all.insert(all.end(), lists[c].begin(), lists[c].end());
      }
      GRMP_TRY(cp->class_tiles.upload(all.data(), all.size(), s));
      GRMP_CUDA(cudaStreamSynchronize(s));
      break;
    }
    if (nw == 1) return fail(GRMP_EUNSUPPORTED, "column kernels: one group of 32 columns does not fit into shared memory");
  }
  temp.release(); ck1.release(); ck2.release(); ci1.release(); ci2.release(); np64.release();
  GRMP_TRY(cp->geo.alloc((size_t)ncells * v_stride));
  // (4) records
  GRMP_TRY(cp->recs.alloc((size_t)std::max<i64>(npairs_used, 1) * cp->nv));
  PackParams pp{};
  pp.colptr = pat.colptr.p; pp.rowval = pat.rowval.p; pp.pairbeg = dg.segptr.p; pp.gcell = dg.gcell.p; pp.gsrc = dg.gsrc.p;
  pp.colperm = cp->colperm.p; pp.pos_recbeg = cp->pos_recbeg.p;
  pp.dofsR = er.celldofs; pp.ndR = er.nd; pp.orient = p.g.orient; pp.regions = p.g.regions; pp.reg = p.reg;
  pp.ncells = ncells; pp.ncols_used = ncols_used;
  pp.nrow = v_nrow; pp.row_bdm3 = (cp->row.kind == 1 && cp->row.nds == 16); pp.col_bdm3 = (cp->col.kind == 1 && cp->col.nds == 16);
  pp.nw = cp->nw; pp.nv = cp->nv; pp.recs = cp->recs.p; pp.err = flags.p;
  pack_records<<<nblk(cp->ngroups * 32, 128), 128, 0, s>>>(pp);
  GRMP_CUDA(cudaGetLastError());
  GRMP_CUDA(cudaMemcpyAsync(hflags, flags.p, 4, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  if (hflags[0]) return fail(GRMP_EUNSUPPORTED, "column kernels: record build failed");
  if (getenv("GRMP_VERBOSE"))
  {
    fprintf(stderr, "[grmp columns] nw %d variant %d (1000+: closed form) tiles %lld pairs %lld geometry record %d B pair record %d B; classes:", cp->nw,
            cf ? 1000 + cp->cf_variant : cp->variant, (long long)cp->ntiles, (long long)cp->npairs, 8 * v_stride, 16 * cp->nv);
    for (auto& c : cp->classes) fprintf(stderr, " %lld tiles <= %d B;", (long long)c.count, c.smem_bytes);
    fprintf(stderr, "\n");
  }
  cp->built = true;
  return GRMP_OK;
}


int colpath_numeric(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, ColPath& cp, double* nzval) {
  if (!cp.built) return fail(GRMP_ESTATE, "column kernels: records not built");
  if (cp.ntiles == 0) return GRMP_OK;
  cudaStream_t s = ctx->stream;
  const bool cf = cp.cf_variant >= 0;
  const Variant& V = VARIANTS[cf ? 0 : cp.variant];
  if (!cf && g_const_owner != cp.uid) {
    GRMP_CUDA(cudaMemcpyToSymbolAsync(c_tabR, cp.tabR.data(), cp.tabR.size() * 8, 0, cudaMemcpyHostToDevice, s));
    GRMP_CUDA(cudaMemcpyToSymbolAsync(c_wq, cp.wq.data(), cp.wq.size() * 8, 0, cudaMemcpyHostToDevice, s));
    g_const_owner = cp.uid;
  }
  ColParams cpar{};
  cpar.g = p.g; cpar.recs = cp.recs.p; cpar.pos_np = cp.pos_np.p; cpar.pos_len = cp.pos_len.p; cpar.pos_start = cp.pos_start.p;
  cpar.pos_recbeg = cp.pos_recbeg.p; cpar.geo = cp.geo.p; cpar.tabC = cp.tabC.p;
  cpar.factor = p.factor; cpar.act_p[0] = p.act_p[0]; cpar.act_p[1] = p.act_p[1]; cpar.nzval = nzval;
  cpar.ncols_used = cp.ncols_used; cpar.ngroups = cp.ngroups; cpar.nw = cp.nw; cpar.nq = cp.nq;
  // update_trafo! / mapderiv! / coefficient data of every cell, once per assembly
  if (cf) GRMP_TRY(CFVARIANTS[cp.cf_variant].geo(p.g, p.act_p, cp.geo.p, s)); else GRMP_TRY(V.geo(p.g, cp.geo.p, s));
  for (const auto& c : cp.classes) {     // big tiles first: they have the fewest CTAs per SM and would otherwise be the tail
    cpar.tile_list = cp.class_tiles.p + c.first;
    if (cf) GRMP_TRY(CFVARIANTS[cp.cf_variant].launch(cpar, (int)c.count, 32 * cp.nw, c.smem_bytes, s));
    else GRMP_TRY(V.launch(cpar, (int)c.count, 32 * cp.nw, c.smem_bytes, s));
  }
  return GRMP_OK;
}


// ==== LinearForm gather kernels ===============================================================================================
namespace {

struct LfParams {
  GridView g;
  const u32* recs;
  const u32* dofperm;
  const unsigned short* pos_np;
  const i64* pos_recbeg;
  const u32* tile_cellptr;
  const u32* tile_cells;
  const u32* tile_list;
  const double* tabC;
  const double* wq;
  const double* fdata;
  double* b;
  double factor;
  i64 ndofs, ngroups;
  int nw, nq, fsrc;
};

// b[dof] += (sum_q (sum_k f_k(x_q) cv[k, l, q]) w_q) * (factor * |T|), cells ascending (linearform.jl:181-220).  One thread owns one
// dof: its value is accumulated in a register and written once; no atomics, order fixed.
template <class Ev>
__global__ void __launch_bounds__(256) lf_kernel(const LfParams p) {
  constexpr int STRIDE = (2 + Ev::CACHE_N + 1) & ~1;       // [0] |T|, [1] cell id, then the evaluator's data
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthr = blockDim.x;
  const int nq = p.nq;
  const int ntab = Ev::NAS * nq * CT_PAD;
  double* const sCt = sm;
  double* const sW = sm + ((ntab + 1) & ~1);
  double* const cache = sW + ((nq + 1) & ~1);
  const i64 tile = p.tile_list[blockIdx.x];
  const u32 c0 = p.tile_cellptr[tile], nct = p.tile_cellptr[tile + 1] - c0;
  for (int i = tid; i < ntab; i += nthr) sCt[i] = p.tabC[i];
  for (int i = tid; i < nq; i += nthr) sW[i] = p.wq[i];
  for (u32 t = tid; t < nct; t += nthr) {
    const i64 cell = p.tile_cells[c0 + t];
    double* cr = cache + (size_t)t * STRIDE;
    CellGeo<Ev::ED> T;
    cell_geo<Ev::ED>(p.g, cell, T);
    cr[0] = p.g.vol[cell];
    cr[1] = __longlong_as_double((long long)cell);
    Ev::build_cache(p.g, cell, T, cr + 2);
  }
  __syncthreads();
  const i64 grp = tile * p.nw + warp;
  if (grp >= p.ngroups) return;
  const i64 pos = grp * 32 + lane;
  const bool has = pos < p.ndofs;
  const u32 np = has ? p.pos_np[pos] : 0u;
  const i64 recbase = p.pos_recbeg[grp * 32];
  const u32 maxnp = __reduce_max_sync(0xffffffffu, np);
  const u32 lt = (1u << lane) - 1u;
  u32 rbase = 0;
  const i64 dof = has ? (i64)p.dofperm[pos] : 0;
  double acc = has ? p.b[dof] : 0.0;
  for (u32 k = 0; k < maxnp; k++) {
    const u32 bal = __ballot_sync(0xffffffffu, k < np);
    const u32 idx = rbase + __popc(bal & lt);
    rbase += __popc(bal);
    if (k >= np) continue;
    const u32 rec = __ldg(p.recs + recbase + idx);
    const u32 lc = (rec >> 16) & 255u;
    if (lc == 255u) continue;
    const double* cr = cache + (size_t)(rec & 0xffffu) * STRIDE;
    typename Ev::Regs R;
    Ev::load(cr + 2, R);
    const i64 cell = __double_as_longlong(cr[1]);
    double lb = 0.0;
    typename Ev::Prep PC;
    Ev::prep(R, (int)lc, PC);
#pragma unroll 1
    for (int q = 0; q < nq; q++) {
      double Y[Ev::RD];
      Ev::col_eval(R, PC, sCt, nq, q, Y);
      double t = 0.0;
      if (p.fsrc == GRMP_F_NONE) {
#pragma unroll
        for (int i = 0; i < Ev::RD; i++) t += Y[i];
      } else if (p.fsrc == GRMP_F_CONST) {
#pragma unroll
        for (int i = 0; i < Ev::RD; i++) t = fma(__ldg(p.fdata + i), Y[i], t);
      } else {
        const double* f = p.fdata + ((size_t)cell * nq + q) * Ev::RD;
#pragma unroll
        for (int i = 0; i < Ev::RD; i++) t = fma(__ldg(f + i), Y[i], t);
      }
      lb = fma(t, sW[q], lb);
    }
    acc += lb * (p.factor * cr[0]);
  }
  if (has) p.b[dof] = acc;
}

__global__ void lf_pack_records(const u32* dofperm, const i64* pairbeg, const i64* pos_recbeg, const u32* gcell, const u32* gsrc,
                                const i32* orient, const i32* regions, RegionFilter reg, const u32* tile_cellptr, const u32* tile_cells,
                                i64 ncells, i64 ndofs, int nw, int col_bdm3, u32* recs, int* err) {
  const i64 grp = (blockIdx.x * (i64)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (grp * 32 >= ndofs) return;
  const i64 pos = grp * 32 + lane;
  const bool has = pos < ndofs;
  const i64 j = has ? (i64)dofperm[pos] : 0;
  const i64 kb = has ? pairbeg[j] : 0;
  const u32 np = has ? (u32)(pairbeg[j + 1] - kb) : 0u;
  const i64 recbase = pos_recbeg[grp * 32];
  const i64 tile = grp / nw;
  const u32 tc0 = tile_cellptr[tile], tc1 = tile_cellptr[tile + 1];
  const u32 maxnp = __reduce_max_sync(0xffffffffu, np);
  const u32 lt = (1u << lane) - 1u;
  u32 rbase = 0;
  for (u32 k = 0; k < maxnp; k++) {
    const u32 bal = __ballot_sync(0xffffffffu, k < np);
    const u32 idx = rbase + __popc(bal & lt);
    rbase += __popc(bal);
    if (k >= np) continue;
    const i64 cell = gcell[kb + k];
    int lc = (int)(gsrc[kb + k] / (u32)ncells);
    bool active = true;
    if (reg.n > 0) {
      active = false;
      if (regions) for (int r = 0; r < reg.n; r++) active = active || regions[cell] == reg.r[r];
    }
    if (col_bdm3) lc = bdm3_ref_of_local(orient + cell * 4, lc);
    u32 lo = tc0, hi = tc1;
    while (lo < hi) { const u32 mid = (lo + hi) >> 1; if (tile_cells[mid] < (u32)cell) lo = mid + 1; else hi = mid; }
    if (lo >= tc1 || tile_cells[lo] != (u32)cell || lo - tc0 > 65535u) atomicExch(err, 2);
    recs[recbase + idx] = (lo - tc0) | ((active ? (u32)lc : 255u) << 16);
  }
}

typedef int (*LfLaunchFn)(const LfParams&, int nblocks, int nthreads, int smem, cudaStream_t);
struct LfVariant {
  bool (*match)(const ColEvalDesc&);
  LfLaunchFn launch;
  int stride, nas;
};
template <class Ev> struct LfVariantImpl {
  static bool match(const ColEvalDesc& d) { return ev_matches<Ev>(d); }
  static int launch(const LfParams& p, int nblocks, int nthreads, int smem, cudaStream_t s) {
    static int attr_set = 0;
    if (smem > attr_set) {
      GRMP_CUDA(cudaFuncSetAttribute(lf_kernel<Ev>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_set = smem;
    }
    lf_kernel<Ev><<<nblocks, nthreads, smem, s>>>(p);
    GRMP_CUDA(cudaGetLastError());
    return GRMP_OK;
  }
};
#define GRMP_LFV(...) {&LfVariantImpl<__VA_ARGS__>::match, &LfVariantImpl<__VA_ARGS__>::launch, (2 + __VA_ARGS__::CACHE_N + 1) & ~1, __VA_ARGS__::NAS},
#define GRMP_LF_H1(ED, NC, NDS, NB) GRMP_LFV(H1Ev<ED, NC, NDS, NB, GRMP_OP_ID>) GRMP_LFV(H1Ev<ED, NC, NDS, NB, GRMP_OP_GRAD>)
const LfVariant LFVARIANTS[] = {
    GRMP_LF_H1(2, 1, 3, 0) GRMP_LF_H1(2, 2, 3, 0) GRMP_LF_H1(2, 1, 6, 0) GRMP_LF_H1(2, 2, 6, 0) GRMP_LF_H1(2, 2, 3, 3) GRMP_LFV(H1Ev<2, 1, 1, 0, GRMP_OP_ID>)
    GRMP_LF_H1(3, 1, 4, 0) GRMP_LF_H1(3, 3, 4, 0) GRMP_LF_H1(3, 1, 10, 0) GRMP_LF_H1(3, 3, 10, 0) GRMP_LF_H1(3, 3, 4, 4) GRMP_LFV(H1Ev<3, 1, 1, 0, GRMP_OP_ID>)
    GRMP_LFV(H1Ev<2, 2, 3, 0, GRMP_OP_DIV>) GRMP_LFV(H1Ev<2, 2, 6, 0, GRMP_OP_DIV>) GRMP_LFV(H1Ev<2, 2, 3, 3, GRMP_OP_DIV>)
    GRMP_LFV(H1Ev<3, 3, 4, 0, GRMP_OP_DIV>) GRMP_LFV(H1Ev<3, 3, 10, 0, GRMP_OP_DIV>) GRMP_LFV(H1Ev<3, 3, 4, 4, GRMP_OP_DIV>)
    GRMP_LFV(HdivEv<2, 3, GRMP_OP_ID>) GRMP_LFV(HdivEv<2, 6, GRMP_OP_ID>) GRMP_LFV(HdivEv<3, 4, GRMP_OP_ID>) GRMP_LFV(HdivEv<3, 16, GRMP_OP_ID>)
    GRMP_LFV(HdivEv<2, 3, GRMP_OP_DIV>) GRMP_LFV(HdivEv<2, 6, GRMP_OP_DIV>) GRMP_LFV(HdivEv<3, 4, GRMP_OP_DIV>) GRMP_LFV(HdivEv<3, 16, GRMP_OP_DIV>)
    GRMP_LFV(ReconEv2D<3>) GRMP_LFV(ReconEv2D<6>)};
constexpr int NLFVARIANTS = sizeof(LFVARIANTS) / sizeof(LFVARIANTS[0]);

}  // namespace

bool lfpath_applicable(const EvalView& e, int edim, int nq, LfPath* lp) {
  if (describe(e, edim, &lp->ev)) return false;
  if (nq > 256) return false;
  lp->variant = -1;
  for (int v = 0; v < NLFVARIANTS; v++)
    if (LFVARIANTS[v].match(lp->ev)) { lp->variant = v; break; }
  lp->nq = nq;
  return lp->variant >= 0;
}

int lfpath_build(grmp_ctx* ctx, const GridView& g, const EvalView& e, const RegionFilter& reg, const std::vector<double>& w,
                 const std::vector<double>& vals, const std::vector<double>& derivs, i64 ndofs, LfPath* lp) {
  cudaStream_t s = ctx->stream;
  lp->built = false;
  const LfVariant& V = LFVARIANTS[lp->variant];
  const int nq = lp->nq;
  const i64 ncells = g.ncells;
  {
    std::vector<double> T;
    int nas = 0;
    GRMP_TRY(make_tables(lp->ev, nq, vals, derivs, e.tab_nd, e.tab_nc, &T, &nas));
    const int nsf = lp->ev.nds + lp->ev.nbub;
    std::vector<double> tc((size_t)nas * nq * CT_PAD, 0.0);
    for (int a = 0; a < nas; a++) for (int q = 0; q < nq; q++) for (int sI = 0; sI < nsf; sI++)
      tc[((size_t)a * nq + q) * CT_PAD + sI] = T[((size_t)sI * nas + a) * nq + q];
    GRMP_TRY(lp->tabC.upload(tc.data(), tc.size(), s));
    GRMP_TRY(lp->wq.upload(w.data(), w.size(), s));
  }
  lp->ndofs = ndofs;
  lp->ngroups = (ndofs + 31) / 32;
  lp->ntiles = 0;
  if (ndofs == 0 || ncells == 0) { lp->built = true; return GRMP_OK; }
  DofGather dg;
  GRMP_TRY(build_dofgather(s, e.celldofs, ncells, e.nd, ndofs, &dg));
  lp->npairs = dg.ncontrib;
  const i64 npairs = dg.ncontrib;
  DevBuf<int> flags;
  GRMP_TRY(flags.alloc(4));
  GRMP_CUDA(cudaMemsetAsync(flags.p, 0, 16, s));
  DevBuf<u64> ck1, ck2, keys, keys2, uniq; DevBuf<u32> ci1; DevBuf<unsigned char> temp; DevBuf<i64> np64, nuniq_d, dummy_start;
  DevBuf<unsigned char> dummy_len;
  GRMP_TRY(ck1.alloc(ndofs)); GRMP_TRY(ck2.alloc(ndofs)); GRMP_TRY(ci1.alloc(ndofs)); GRMP_TRY(lp->dofperm.alloc(ndofs));
  order_keys<<<nblk(ndofs), 256, 0, s>>>(nullptr, dg.segptr.p, dg.gcell.p, ndofs, ck1.p, ci1.p, flags.p);
  GRMP_CUDA(cudaGetLastError());
  {
    cub::DoubleBuffer<u64> dk(ck1.p, ck2.p);
    cub::DoubleBuffer<u32> dv(ci1.p, lp->dofperm.p);
    size_t tb = 0;
    GRMP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, ndofs, 0, 64, s));
    GRMP_TRY(temp.alloc(tb));
    GRMP_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, tb, dk, dv, ndofs, 0, 64, s));
    if (dv.Current() != lp->dofperm.p) GRMP_CUDA(cudaMemcpyAsync(lp->dofperm.p, dv.Current(), (size_t)ndofs * 4, cudaMemcpyDeviceToDevice, s));
  }
  GRMP_TRY(lp->pos_np.alloc(ndofs)); GRMP_TRY(lp->pos_recbeg.alloc(ndofs + 1)); GRMP_TRY(np64.alloc(ndofs + 1));
  GRMP_TRY(dummy_len.alloc(ndofs)); GRMP_TRY(dummy_start.alloc(ndofs));
  // colptr of a pseudo pattern with one entry per dof: reuse pos_arrays with the pair offsets as "colptr"
  pos_arrays<<<nblk(ndofs), 256, 0, s>>>(lp->dofperm.p, dg.segptr.p, dg.segptr.p, ndofs, lp->pos_np.p, dummy_len.p, dummy_start.p, np64.p);
  GRMP_CUDA(cudaGetLastError());
  GRMP_CUDA(cudaMemsetAsync(np64.p + ndofs, 0, 8, s));
  {
    size_t tb = 0;
    GRMP_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, np64.p, lp->pos_recbeg.p, ndofs + 1, s));
    if (tb > temp.n) GRMP_TRY(temp.alloc(tb));
    GRMP_CUDA(cub::DeviceScan::ExclusiveSum(temp.p, tb, np64.p, lp->pos_recbeg.p, ndofs + 1, s));
  }
  GRMP_TRY(keys.alloc(std::max<i64>(npairs, 1))); GRMP_TRY(keys2.alloc(std::max<i64>(npairs, 1))); GRMP_TRY(uniq.alloc(std::max<i64>(npairs, 1)));
  GRMP_TRY(nuniq_d.alloc(1));
  int nw = 4;
  int hflags[4] = {0, 0, 0, 0};
  for (;; nw >>= 1) {
    const int cpt = 32 * nw;
    const i64 ntiles = (ndofs + cpt - 1) / cpt;
    tile_keys<<<nblk(ndofs), 256, 0, s>>>(lp->dofperm.p, dg.segptr.p, lp->pos_recbeg.p, dg.gcell.p, ndofs, cpt, keys.p);
    GRMP_CUDA(cudaGetLastError());
    int end_bit = 33;
    while (end_bit < 64 && ((u64)ntiles >> (end_bit - 32)) != 0) end_bit++;
    size_t tb = 0, tb2 = 0;
    GRMP_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, keys.p, keys2.p, npairs, 0, end_bit, s));
    GRMP_CUDA(cub::DeviceSelect::Unique(nullptr, tb2, keys2.p, uniq.p, nuniq_d.p, npairs, s));
    if (std::max(tb, tb2) > temp.n) GRMP_TRY(temp.alloc(std::max(tb, tb2)));
    GRMP_CUDA(cub::DeviceRadixSort::SortKeys(temp.p, tb, keys.p, keys2.p, npairs, 0, end_bit, s));
    GRMP_CUDA(cub::DeviceSelect::Unique(temp.p, tb2, keys2.p, uniq.p, nuniq_d.p, npairs, s));
    i64 nuniq = 0;
    GRMP_CUDA(cudaMemcpyAsync(&nuniq, nuniq_d.p, 8, cudaMemcpyDeviceToHost, s));
    GRMP_CUDA(cudaStreamSynchronize(s));
    GRMP_TRY(lp->tile_cellptr.alloc(ntiles + 1));
    GRMP_TRY(lp->tile_cells.alloc(std::max<i64>(nuniq, 1)));
    GRMP_CUDA(cudaMemsetAsync(flags.p + 2, 0, 4, s));
    tile_ptr<<<nblk(ntiles + 1), 256, 0, s>>>(uniq.p, nuniq, ntiles, lp->tile_cellptr.p, flags.p + 2);
    low32<<<nblk(nuniq), 256, 0, s>>>(uniq.p, nuniq, lp->tile_cells.p);
    GRMP_CUDA(cudaGetLastError());
    std::vector<u32> tcp(ntiles + 1);
    GRMP_CUDA(cudaMemcpyAsync(tcp.data(), lp->tile_cellptr.p, (size_t)(ntiles + 1) * 4, cudaMemcpyDeviceToHost, s));
    GRMP_CUDA(cudaMemcpyAsync(hflags, flags.p, 16, cudaMemcpyDeviceToHost, s));
    GRMP_CUDA(cudaStreamSynchronize(s));
    if (hflags[0]) return fail(GRMP_EUNSUPPORTED, "linear form kernels: a dof lies in more than 65535 cells");
    const i64 fixed = 8 * ((((i64)V.nas * nq * CT_PAD + 1) & ~1) + ((nq + 1) & ~1));
    const i64 need_max = fixed + (i64)hflags[2] * V.stride * 8;
    if (hflags[2] <= 65535 && need_max <= 200 * 1024) {
      lp->nw = nw; lp->ntiles = ntiles;
      const int caps[6] = {6 * 1024, 13 * 1024, 27 * 1024, 55 * 1024, 112 * 1024, 200 * 1024};
      std::vector<std::vector<u32>> lists(6);
      std::vector<int> mx(6, 0);
      for (i64 t = 0; t < ntiles; t++) {
        const int need = (int)(fixed + (i64)(tcp[t + 1] - tcp[t]) * V.stride * 8);
        int c = 0;
        while (c < 5 && need > caps[c]) c++;
        lists[c].push_back((u32)t);
        mx[c] = std::max(mx[c], need);
      }
      std::vector<u32> all;
      lp->classes.clear();
      for (int c = 5; c >= 0; c--) {
        if (lists[c].empty()) continue;
        lp->classes.push_back(ColPath::TileClass{mx[c], (i64)all.size(), (i64)lists[c].size()});
        all.insert(all.end(), lists[c].begin(), lists[c].end());
      }
      GRMP_TRY(lp->class_tiles.upload(all.data(), all.size(), s));
      GRMP_CUDA(cudaStreamSynchronize(s));
      break;
    }
    if (nw == 1) return fail(GRMP_EUNSUPPORTED, "linear form kernels: one group of 32 dofs does not fit into shared memory");
  }
  GRMP_TRY(lp->recs.alloc(std::max<i64>(npairs, 1)));
  lf_pack_records<<<nblk(lp->ngroups * 32, 128), 128, 0, s>>>(lp->dofperm.p, dg.segptr.p, lp->pos_recbeg.p, dg.gcell.p, dg.gsrc.p, g.orient, g.regions,
                                                             reg, lp->tile_cellptr.p, lp->tile_cells.p, ncells, ndofs, lp->nw,
                                                             (lp->ev.kind == 1 && lp->ev.nds == 16) ? 1 : 0, lp->recs.p, flags.p);
  GRMP_CUDA(cudaGetLastError());
  GRMP_CUDA(cudaMemcpyAsync(hflags, flags.p, 4, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  if (hflags[0]) return fail(GRMP_EUNSUPPORTED, "linear form kernels: record build failed");
  lp->built = true;
  return GRMP_OK;
}

int lfpath_numeric(grmp_ctx* ctx, const GridView& g, LfPath& lp, double factor, int fsrc, const double* fdata, double* b) {
  if (!lp.built) return fail(GRMP_ESTATE, "linear form kernels: records not built");
  if (lp.ntiles == 0) return GRMP_OK;
  const LfVariant& V = LFVARIANTS[lp.variant];
  LfParams p{};
  p.g = g; p.recs = lp.recs.p; p.dofperm = lp.dofperm.p; p.pos_np = lp.pos_np.p; p.pos_recbeg = lp.pos_recbeg.p;
  p.tile_cellptr = lp.tile_cellptr.p; p.tile_cells = lp.tile_cells.p; p.tabC = lp.tabC.p; p.wq = lp.wq.p; p.fdata = fdata; p.b = b;
  p.factor = factor; p.ndofs = lp.ndofs; p.ngroups = lp.ngroups; p.nw = lp.nw; p.nq = lp.nq; p.fsrc = fsrc;
  for (const auto& c : lp.classes) {
    p.tile_list = lp.class_tiles.p + c.first;
    GRMP_TRY(V.launch(p, (int)c.count, 32 * lp.nw, c.smem_bytes, ctx->stream));
  }
  return GRMP_OK;
}

}  // namespace grmp
