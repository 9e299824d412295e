// fastpath_p2tet.cu -- owner-computes numeric kernel for the metric configuration:
// 3D P2 Laplace stiffness (H1P2{1,3}, [Gradient, Gradient], NoAction) on a frozen pattern.
//
// Replaces, for this form, the whole cell loop of assemble! (bilinearform.jl:226-377:
// update_trafo!/mapderiv!, update_basis! Gradient, quadrature contraction, _addnz scatter)
// by ONE kernel in which every stored non-zero is computed and written exactly once:
//
//   * a tile = a contiguous range of CSC columns (<= 128) whose nzval range is staged in
//     shared memory and written back with fully coalesced stores;
//   * the geometry S_ab = factor*|T| grad(lambda_a).grad(lambda_b) of the tile's distinct
//     cells is computed once per tile into shared memory (the affine pullback of
//     feevaluator_h1.jl:61-74 reduced to its 10 invariants);
//   * one thread owns one column j and walks the cells K containing dof j ("pairs"); for
//     each pair it evaluates the whole local column K_loc[:, lj] from S with the exact
//     P2 integrals (the order-2 rule of quadrature.jl:274-284 integrates them exactly) and
//     adds the 10 values into the column's slots through a 1-byte local->nnz map.
//
// No atomics, no inter-block dependencies -> deterministic.  Values agree with the
// reference order of operations to rounding (<= 1e-12 relative, tests/test_gpu_parity.py);
// the *pattern* always comes from the bit-exact symbolic pass.
#include <algorithm>
#include <cmath>

#include "fastpath.cuh"

namespace grmp {

namespace {

constexpr int TPB = 128;             // threads per tile
constexpr int MAX_TILE_COLS = 128;
constexpr int SMEM_BUDGET = 56 * 1024;          // staged tiles: nzval slots + S of the distinct cells
constexpr int SMEM_BUDGET_DIRECT = 56 * 1024;   // direct tiles (vertex columns): nzval slots + 32 B per pair

// local edge e -> (p,q), Tetrahedron3D edges [1 2],[1 3],[1 4],[2 3],[2 4],[3 4] (h1_p2.jl:231-236)
__host__ __device__ inline void edge_nodes(int e, int& p, int& q) {
  const int P[6] = {0, 0, 0, 1, 1, 2}, Q[6] = {1, 2, 3, 2, 3, 3};
  p = P[e]; q = Q[e];
}
__host__ __device__ inline int edge_of(int a, int b) {  // a < b
  return (a == 0) ? (b - 1) : (a == 1 ? b + 1 : 5);
}
__host__ __device__ inline int sidx(int a, int b) {      // index of S_ab in the packed upper triangle
  if (a > b) { int t = a; a = b; b = t; }
  return a * 4 - a * (a - 1) / 2 + (b - a);
}
// canonical vertex permutation of a pair whose column is local dof lj:
// vertex column a -> (a, others ascending); edge column (p,q) -> (p, q, others ascending)
__host__ __device__ inline void canon_perm(int lj, int* pi) {
  int used[4] = {0, 0, 0, 0}, n = 0;
  if (lj < 4) { pi[n++] = lj; used[lj] = 1; }
  else { int p, q; edge_nodes(lj - 4, p, q); pi[n++] = p; pi[n++] = q; used[p] = used[q] = 1; }
  for (int v = 0; v < 4; v++) if (!used[v]) pi[n++] = v;
}
// canonical row r (v_pi0..v_pi3, e(pi0pi1), e(pi0pi2), e(pi0pi3), e(pi1pi2), e(pi1pi3), e(pi2pi3)) -> local dof
__host__ __device__ inline int canon_row(const int* pi, int r) {
  if (r < 4) return pi[r];
  const int A[6] = {0, 0, 0, 1, 1, 2}, B[6] = {1, 2, 3, 2, 3, 3};
  int a = pi[A[r - 4]], b = pi[B[r - 4]];
  return 4 + (a < b ? edge_of(a, b) : edge_of(b, a));
}

struct PackParams {
  const u32* gsrc;        // sorted pairs: lj*ncells + cell
  const u32* gcell;
  const u32* pair_x;      // tile-local cell index (staged tiles) or global cell index (direct tiles)
  const i32* slotmap;     // [ncells*100]
  const i64* colptr;      // 1-based
  const i32* celldofs;
  i64 npairs, ncells;
  uint4* pairs;
};

__global__ void pack_pairs(PackParams p) {
  i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (k >= p.npairs) return;
  const i64 cell = p.gcell[k];
  const int lj = (int)(p.gsrc[k] / (u32)p.ncells);
  const i64 col = p.celldofs[cell * 10 + lj] - 1;
  const i64 base = p.colptr[col] - 1;
  int pi[4];
  canon_perm(lj, pi);
  unsigned char off[12];
  for (int r = 0; r < 10; r++) {
    const int li = canon_row(pi, r);
    const i32 slot = p.slotmap[cell * 100 + li * 10 + lj];
    off[r] = (slot < 0) ? 255 : (unsigned char)(slot - base);
  }
  off[10] = (unsigned char)lj; off[11] = 0;
  uint4 rec;
  rec.x = p.pair_x[k];
  rec.y = off[0] | (off[1] << 8) | (off[2] << 16) | ((u32)off[3] << 24);
  rec.z = off[4] | (off[5] << 8) | (off[6] << 16) | ((u32)off[7] << 24);
  rec.w = off[8] | (off[9] << 8) | (off[10] << 16) | ((u32)off[11] << 24);
  p.pairs[k] = rec;
}

struct TileParams {
  GridView g;
  const i64* colptr;        // 1-based [ncols+1]
  const i64* col_pairbeg;   // [ncols+1]
  const uint4* pairs;
  const i32* tile_colbeg;
  const i32* tile_cellbeg;
  const i32* tile_cells;
  double factor;
  double* nzval;
};

// S_ab = factor * |T| * grad(lambda_a).grad(lambda_b), packed (00,01,02,03,11,12,13,22,23,33)
__device__ __forceinline__ void cell_S(const GridView& g, i64 cell, double factor, double* S) {
  const i32* cn = g.cellnodes + cell * 4;
  const double* x0 = g.coords + (i64)(cn[0] - 1) * 3;
  const double* x1 = g.coords + (i64)(cn[1] - 1) * 3;
  const double* x2 = g.coords + (i64)(cn[2] - 1) * 3;
  const double* x3 = g.coords + (i64)(cn[3] - 1) * 3;
  const double ax = x1[0] - x0[0], ay = x1[1] - x0[1], az = x1[2] - x0[2];
  const double bx = x2[0] - x0[0], by = x2[1] - x0[1], bz = x2[2] - x0[2];
  const double cx = x3[0] - x0[0], cy = x3[1] - x0[1], cz = x3[2] - x0[2];
  // n1 = b x c, n2 = c x a, n3 = a x b : grad(lambda_k) = n_k / det
  const double n1x = by * cz - bz * cy, n1y = bz * cx - bx * cz, n1z = bx * cy - by * cx;
  const double n2x = cy * az - cz * ay, n2y = cz * ax - cx * az, n2z = cx * ay - cy * ax;
  const double n3x = ay * bz - az * by, n3y = az * bx - ax * bz, n3z = ax * by - ay * bx;
  const double n0x = -(n1x + n2x + n3x), n0y = -(n1y + n2y + n3y), n0z = -(n1z + n2z + n3z);
  // |T| / det^2 with det = 6|T| taken from CellVolumes like mapderiv! does: 1 / (36 |T|)
  const double sc = factor / (36.0 * g.vol[cell]);
  S[0] = sc * (n0x * n0x + n0y * n0y + n0z * n0z);
  S[1] = sc * (n0x * n1x + n0y * n1y + n0z * n1z);
  S[2] = sc * (n0x * n2x + n0y * n2y + n0z * n2z);
  S[3] = sc * (n0x * n3x + n0y * n3y + n0z * n3z);
  S[4] = sc * (n1x * n1x + n1y * n1y + n1z * n1z);
  S[5] = sc * (n1x * n2x + n1y * n2y + n1z * n2z);
  S[6] = sc * (n1x * n3x + n1y * n3y + n1z * n3z);
  S[7] = sc * (n2x * n2x + n2y * n2y + n2z * n2z);
  S[8] = sc * (n2x * n3x + n2y * n3y + n2z * n3z);
  S[9] = sc * (n3x * n3x + n3y * n3y + n3z * n3z);
}

// the four values a vertex column a needs, straight from the coordinates (direct tiles):
// S_aa and S_ab for the other three vertices in ascending order
__device__ __forceinline__ void vertex_S(const GridView& g, i64 cell, int a, double factor, double& saa, double& s1, double& s2, double& s3) {
  const i32* cn = g.cellnodes + cell * 4;
  const int4 nd = *reinterpret_cast<const int4*>(cn);
  const double* x0 = g.coords + (i64)(nd.x - 1) * 3;
  const double* x1 = g.coords + (i64)(nd.y - 1) * 3;
  const double* x2 = g.coords + (i64)(nd.z - 1) * 3;
  const double* x3 = g.coords + (i64)(nd.w - 1) * 3;
  const double p0x = x0[0], p0y = x0[1], p0z = x0[2];
  const double ax = x1[0] - p0x, ay = x1[1] - p0y, az = x1[2] - p0z;
  const double bx = x2[0] - p0x, by = x2[1] - p0y, bz = x2[2] - p0z;
  const double cx = x3[0] - p0x, cy = x3[1] - p0y, cz = x3[2] - p0z;
  const double n1x = by * cz - bz * cy, n1y = bz * cx - bx * cz, n1z = bx * cy - by * cx;
  const double n2x = cy * az - cz * ay, n2y = cz * ax - cx * az, n2z = cx * ay - cy * ax;
  const double n3x = ay * bz - az * by, n3y = az * bx - ax * bz, n3z = ax * by - ay * bx;
  const double n0x = -(n1x + n2x + n3x), n0y = -(n1y + n2y + n3y), n0z = -(n1z + n2z + n3z);
  const double nax = a == 0 ? n0x : a == 1 ? n1x : a == 2 ? n2x : n3x;
  const double nay = a == 0 ? n0y : a == 1 ? n1y : a == 2 ? n2y : n3y;
  const double naz = a == 0 ? n0z : a == 1 ? n1z : a == 2 ? n2z : n3z;
  const double sc = factor / (36.0 * g.vol[cell]);
  const double d0 = sc * (nax * n0x + nay * n0y + naz * n0z);
  const double d1 = sc * (nax * n1x + nay * n1y + naz * n1z);
  const double d2 = sc * (nax * n2x + nay * n2y + naz * n2z);
  const double d3 = sc * (nax * n3x + nay * n3y + naz * n3z);
  saa = a == 0 ? d0 : a == 1 ? d1 : a == 2 ? d2 : d3;
  s1 = (a == 0) ? d1 : d0;
  s2 = (a <= 1) ? d2 : d1;
  s3 = (a <= 2) ? d3 : d2;
}

// add the 10 values of one pair into the column's slots; the 10 rows of a pair are distinct slots,
// so all reads are issued before the writes (no dependent read-modify-write chain)
__device__ __forceinline__ void pair_update(double* a, const uint4& rec, const double* v) {
  u32 o[10];
  o[0] = rec.y & 255u; o[1] = (rec.y >> 8) & 255u; o[2] = (rec.y >> 16) & 255u; o[3] = rec.y >> 24;
  o[4] = rec.z & 255u; o[5] = (rec.z >> 8) & 255u; o[6] = (rec.z >> 16) & 255u; o[7] = rec.z >> 24;
  o[8] = rec.w & 255u; o[9] = (rec.w >> 8) & 255u;
  double c[10];
#pragma unroll
  for (int r = 0; r < 10; r++) c[r] = (o[r] != 255u) ? a[o[r]] : 0.0;
#pragma unroll
  for (int r = 0; r < 10; r++)
    if (o[r] != 255u) a[o[r]] = c[r] + v[r];
}

__device__ __forceinline__ void vertex_values(double saa, double s1, double s2, double s3, double* v) {
  const double m = -0.2 * saa;
  v[0] = 0.6 * saa;
  v[1] = -0.2 * s1; v[2] = -0.2 * s2; v[3] = -0.2 * s3;
  v[4] = 0.6 * s1 + m; v[5] = 0.6 * s2 + m; v[6] = 0.6 * s3 + m;
  v[7] = -0.2 * (s1 + s2); v[8] = -0.2 * (s1 + s3); v[9] = -0.2 * (s2 + s3);
}

__device__ __forceinline__ void edge_values(const double* s, const unsigned char* ix, double* v) {
  const double spp = s[ix[0]], sqq = s[ix[1]], spq = s[ix[2]], spr = s[ix[3]], sps = s[ix[4]], sqr = s[ix[5]], sqs = s[ix[6]];
  v[0] = 0.6 * spq - 0.2 * spp;
  v[1] = 0.6 * spq - 0.2 * sqq;
  v[2] = -0.2 * (spr + sqr);
  v[3] = -0.2 * (sps + sqs);
  v[4] = 1.6 * (spp + sqq + spq);
  v[5] = 0.8 * (2.0 * sqr + spq + spr + spp);
  v[6] = 0.8 * (2.0 * sqs + spq + sps + spp);
  v[7] = 0.8 * (2.0 * spr + spq + sqr + sqq);
  v[8] = 0.8 * (2.0 * sps + spq + sqs + sqq);
  v[9] = 0.8 * (spr + sps + sqr + sqs);
}

constexpr int PF = 4;   // pair records fetched per batch (independent 16-byte loads in flight per thread)

__global__ void __launch_bounds__(TPB) p2tet_tile_kernel(const TileParams p) {
  extern __shared__ double sm[];
  __shared__ unsigned char s_sidx[10][8];   // packed-S positions needed by a column of local dof lj
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int c0 = p.tile_colbeg[tile], c1 = p.tile_colbeg[tile + 1];
  const int cb = p.tile_cellbeg[tile], nct = p.tile_cellbeg[tile + 1] - cb;
  const i64 g0 = p.colptr[c0] - 1, g1 = p.colptr[c1] - 1;
  const int nnz_t = (int)(g1 - g0);
  const bool staged = nct > 0;          // staged: S of the distinct cells in smem; direct: per-pair values in smem
  double* acc = sm;
  double* S = sm + nnz_t;
  if (tid < 10) {
    int pi[4];
    canon_perm(tid, pi);
    if (tid < 4) {   // vertex column a: S_aa, S_ab1, S_ab2, S_ab3
      for (int k = 0; k < 4; k++) s_sidx[tid][k] = (unsigned char)sidx(pi[0], pi[k]);
      for (int k = 4; k < 8; k++) s_sidx[tid][k] = 0;
    } else {         // edge column (p,q | r,s): S_pp, S_qq, S_pq, S_pr, S_ps, S_qr, S_qs
      s_sidx[tid][0] = (unsigned char)sidx(pi[0], pi[0]);
      s_sidx[tid][1] = (unsigned char)sidx(pi[1], pi[1]);
      s_sidx[tid][2] = (unsigned char)sidx(pi[0], pi[1]);
      s_sidx[tid][3] = (unsigned char)sidx(pi[0], pi[2]);
      s_sidx[tid][4] = (unsigned char)sidx(pi[0], pi[3]);
      s_sidx[tid][5] = (unsigned char)sidx(pi[1], pi[2]);
      s_sidx[tid][6] = (unsigned char)sidx(pi[1], pi[3]);
      s_sidx[tid][7] = 0;
    }
  }
  // this thread's column (tiles hold at most TPB columns) and its first batch of pair records:
  // issued before the geometry phase so that the loads overlap it
  const int col = c0 + tid;
  const bool has_col = col < c1;
  i64 kb = 0, ke = 0;
  int abase = 0;
  if (has_col) {
    kb = p.col_pairbeg[col]; ke = p.col_pairbeg[col + 1];
    abase = (int)(p.colptr[col] - 1 - g0);
  }
  uint4 rec[PF];
#pragma unroll
  for (int j = 0; j < PF; j++) rec[j] = (kb + j < ke) ? p.pairs[kb + j] : make_uint4(0, 0, 0, 0);
  for (int i = tid; i < nnz_t; i += TPB) acc[i] = 0.0;

  if (staged) {
    // ---- geometry of the tile's distinct cells, two cells per thread and round (independent loads) ----
    for (int i = tid; i < nct; i += 2 * TPB) {
      const int i2 = i + TPB;
      const i64 cA = p.tile_cells[cb + i];
      const i64 cB = (i2 < nct) ? p.tile_cells[cb + i2] : cA;
      double sA[10], sB[10];
      cell_S(p.g, cA, p.factor, sA);
      cell_S(p.g, cB, p.factor, sB);
#pragma unroll
      for (int k = 0; k < 10; k++) S[i * 10 + k] = sA[k];
      if (i2 < nct) {
#pragma unroll
        for (int k = 0; k < 10; k++) S[i2 * 10 + k] = sB[k];
      }
    }
    __syncthreads();
    if (has_col) {
      double* a = acc + abase;
      for (i64 k = kb; k < ke; k += PF) {
        uint4 cur[PF];
#pragma unroll
        for (int j = 0; j < PF; j++) cur[j] = rec[j];
#pragma unroll
        for (int j = 0; j < PF; j++) rec[j] = (k + PF + j < ke) ? p.pairs[k + PF + j] : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < PF; j++) {
          if (k + j < ke) {
            const int lj = (int)((cur[j].w >> 16) & 255u);
            const double* s = S + cur[j].x * 10;
            const unsigned char* ix = s_sidx[lj];
            double v[10];
            if (lj < 4) vertex_values(s[ix[0]], s[ix[1]], s[ix[2]], s[ix[3]], v);
            else edge_values(s, ix, v);
            pair_update(a, cur[j], v);
          }
        }
      }
    }
  } else {
    // ---- direct tile (vertex columns): phase 1, one thread per PAIR computes its four S values from the
    //      coordinates (PF pairs per thread in flight); phase 2, one thread per column accumulates ----
    const i64 kt0 = p.col_pairbeg[c0], kt1 = p.col_pairbeg[c1];
    const int npt = (int)(kt1 - kt0);
    for (int q0 = 0; q0 < npt; q0 += PF * TPB) {
      uint4 r4[PF];
#pragma unroll
      for (int j = 0; j < PF; j++) {
        const int q = q0 + j * TPB + tid;
        r4[j] = (q < npt) ? p.pairs[kt0 + q] : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int j = 0; j < PF; j++) {
        const int q = q0 + j * TPB + tid;
        if (q < npt) {
          double saa, s1, s2, s3;
          vertex_S(p.g, (i64)r4[j].x, (int)((r4[j].w >> 16) & 255u), p.factor, saa, s1, s2, s3);
          double* st = S + (size_t)q * 4;
          st[0] = saa; st[1] = s1; st[2] = s2; st[3] = s3;
        }
      }
    }
    __syncthreads();
    if (has_col) {
      double* a = acc + abase;
      for (i64 k = kb; k < ke; k += PF) {
        uint4 cur[PF];
#pragma unroll
        for (int j = 0; j < PF; j++) cur[j] = rec[j];
#pragma unroll
        for (int j = 0; j < PF; j++) rec[j] = (k + PF + j < ke) ? p.pairs[k + PF + j] : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < PF; j++) {
          if (k + j < ke) {
            const double* st = S + (size_t)(k + j - kt0) * 4;
            double v[10];
            vertex_values(st[0], st[1], st[2], st[3], v);
            pair_update(a, cur[j], v);
          }
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < nnz_t; i += TPB) p.nzval[g0 + i] = acc[i];
}

// closed-form local stiffness of the unit reference tetrahedron, used to verify that the
// caller's tables describe the standard P2 basis (src/fedefs/h1_p2.jl:223-239)
void reference_local_closed_form(double K[10][10]) {
  double g[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  double S[4][4];
  for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) S[a][b] = (g[a][0] * g[b][0] + g[a][1] * g[b][1] + g[a][2] * g[b][2]) / 6.0;
  auto mm = [](int a, int b) { return a == b ? 0.15 : -0.05; };
  for (int i = 0; i < 10; i++) for (int j = 0; j < 10; j++) {
    double v;
    if (i < 4 && j < 4) v = S[i][j] * (i == j ? 0.6 : -0.2);
    else if (i < 4 || j < 4) {
      int a = i < 4 ? i : j, e = (i < 4 ? j : i) - 4, b, c;
      edge_nodes(e, b, c);
      v = 4 * (S[a][c] * mm(a, b) + S[a][b] * mm(a, c));
    } else {
      int a, b, c, d;
      edge_nodes(i - 4, a, b); edge_nodes(j - 4, c, d);
      v = 0.8 * ((1 + (a == c)) * S[b][d] + (1 + (a == d)) * S[b][c] + (1 + (b == c)) * S[a][d] + (1 + (b == d)) * S[a][c]);
    }
    K[i][j] = v;
  }
}

}  // namespace

bool fast_p2tet_applicable(const BlfLocalParams& p) {
  return p.g.dim == 3 && p.same_eval && p.e1.fam == FAM_H1 && p.e1.op == GRMP_OP_GRAD && p.e1.ncomp == 1 && p.e1.nd == 10 &&
         p.e1.tab_nd == 10 && p.action == GRMP_ACT_NONE && (p.apt == GRMP_APT_SYMMETRIC || (p.apt == GRMP_APT_BILINEARFORM && !p.transposed));
}

int fast_p2tet_build(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, const std::vector<double>& w,
                     const std::vector<double>& derivs, i64 ncols_owned, FastP2Tet* out) {
  cudaStream_t s = ctx->stream;
  const i64 ncells = p.g.ncells, ncols = pat.ncols;
  const i64 ncols_tiled = (ncols_owned >= 0 && ncols_owned < ncols) ? ncols_owned : ncols;   // halo columns belong to another rank
  out->ntiles = 0;
  // (0) the caller's tables must be the standard P2 basis integrated exactly
  {
    const int nq = p.nq;
    if ((int)w.size() != nq || derivs.size() != (size_t)nq * 3 * 10) return fail(GRMP_EINVAL, "fast path: table shape");
    double K[10][10];
    reference_local_closed_form(K);
    for (int i = 0; i < 10; i++) for (int j = 0; j < 10; j++) {
      double v = 0;
      for (int q = 0; q < nq; q++) for (int k = 0; k < 3; k++) v += w[q] * derivs[((size_t)q * 3 + k) * 10 + i] * derivs[((size_t)q * 3 + k) * 10 + j];
      if (std::fabs(v / 6.0 - K[i][j]) > 1e-13) return fail(GRMP_EUNSUPPORTED, "fast path: tables are not the standard P2 basis / exact rule");
    }
  }
  // (1) pairs (cell, lj) sorted by column, cells ascending
  DofGather dg;
  GRMP_TRY(build_dofgather(s, p.e1.celldofs, ncells, 10, ncols, &dg));
  const i64 npairs = dg.ncontrib;
  std::vector<u32> h_cell(npairs), h_src(npairs);
  std::vector<i64> h_pairbeg(ncols + 1), h_colptr(ncols + 1);
  GRMP_CUDA(cudaMemcpyAsync(h_cell.data(), dg.gcell.p, npairs * 4, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaMemcpyAsync(h_src.data(), dg.gsrc.p, npairs * 4, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaMemcpyAsync(h_pairbeg.data(), dg.segptr.p, (ncols + 1) * 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaMemcpyAsync(h_colptr.data(), pat.colptr.p, (ncols + 1) * 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  // (2) greedy tiles over the column range (host; one pass over the pairs)
  std::vector<i32> tile_colbeg{0}, tile_cellbeg{0}, tile_cells;
  std::vector<u32> pair_x(npairs);
  int cur_direct = -1;   // mode of the open tile: 1 = direct (vertex columns, geometry per pair), 0 = staged
  std::vector<i32> mark(ncells, -1), local_of(ncells, 0);
  tile_cells.reserve((size_t)ncells * 4);
  int cur_tile = 0, cur_cols = 0, cur_cells = 0;
  i64 cur_nnz = 0, cur_pairs = 0;
  int max_smem = 0, max_cells = 0;
  for (i64 j = 0; j < ncols_tiled; j++) {
    const i64 len = h_colptr[j + 1] - h_colptr[j];
    if (len > 254) return fail(GRMP_EUNSUPPORTED, "fast path: a column has more than 254 entries");
    const bool has_pairs = h_pairbeg[j + 1] > h_pairbeg[j];
    const int direct = (has_pairs && (h_src[h_pairbeg[j]] / (u32)ncells) < 4) ? 1 : 0;   // vertex dof of the P2 element
    for (int attempt = 0; attempt < 2; attempt++) {
      int fresh = 0;
      if (!direct)
        for (i64 k = h_pairbeg[j]; k < h_pairbeg[j + 1]; k++) if (mark[h_cell[k]] != cur_tile) fresh++;
      const i64 npj = h_pairbeg[j + 1] - h_pairbeg[j];
      const i64 need = direct ? 8 * (cur_nnz + len) + 32 * (cur_pairs + npj) : 8 * (cur_nnz + len) + 80 * (i64)(cur_cells + fresh);
      const i64 budget = direct ? SMEM_BUDGET_DIRECT : SMEM_BUDGET;
      if (cur_cols > 0 && (cur_cols + 1 > MAX_TILE_COLS || need > budget || direct != cur_direct)) {
        tile_colbeg.push_back((i32)j); tile_cellbeg.push_back((i32)tile_cells.size());
        max_smem = std::max<i64>(max_smem, 8 * cur_nnz + (cur_direct ? 32 * cur_pairs : 80 * (i64)cur_cells)); max_cells = std::max(max_cells, cur_cells);
        cur_tile++; cur_cols = 0; cur_cells = 0; cur_nnz = 0; cur_pairs = 0;
        continue;   // re-evaluate the column in the fresh tile
      }
      cur_direct = direct;
      for (i64 k = h_pairbeg[j]; k < h_pairbeg[j + 1]; k++) {
        const u32 c = h_cell[k];
        if (direct) { pair_x[k] = c; continue; }
        if (mark[c] != cur_tile) { mark[c] = cur_tile; local_of[c] = cur_cells++; tile_cells.push_back((i32)c); }
        pair_x[k] = (u32)local_of[c];
      }
      cur_cols++; cur_nnz += len; cur_pairs += npj;
      break;
    }
  }
  if (cur_cols > 0 || ncols_tiled == 0) {
    tile_colbeg.push_back((i32)ncols_tiled); tile_cellbeg.push_back((i32)tile_cells.size());
    max_smem = std::max<i64>(max_smem, 8 * cur_nnz + (cur_direct == 1 ? 32 * cur_pairs : 80 * (i64)cur_cells)); max_cells = std::max(max_cells, cur_cells);
  }
  if (max_smem > 200 * 1024) return fail(GRMP_EUNSUPPORTED, "fast path: a single column exceeds the shared-memory tile");
  out->ntiles = (int)tile_colbeg.size() - 1;
  out->npairs = npairs; out->smem_bytes = max_smem; out->max_tile_cells = max_cells;
  GRMP_TRY(out->tile_colbeg.upload(tile_colbeg.data(), tile_colbeg.size(), s));
  GRMP_TRY(out->tile_cellbeg.upload(tile_cellbeg.data(), tile_cellbeg.size(), s));
  GRMP_TRY(out->tile_cells.upload(tile_cells.data(), tile_cells.size(), s));
  DevBuf<u32> d_local;
  GRMP_TRY(d_local.upload(pair_x.data(), npairs, s));
  // (3) pack the 16-byte pair records on the device
  GRMP_TRY(out->pairs.alloc(npairs));
  GRMP_TRY(out->col_pairbeg.alloc(ncols + 1));
  GRMP_CUDA(cudaMemcpyAsync(out->col_pairbeg.p, dg.segptr.p, (ncols + 1) * 8, cudaMemcpyDeviceToDevice, s));
  if (npairs) {
    PackParams pp{dg.gsrc.p, dg.gcell.p, d_local.p, pat.slotmap.p, pat.colptr.p, p.e1.celldofs, npairs, ncells, out->pairs.p};
    pack_pairs<<<(unsigned)((npairs + 255) / 256), 256, 0, s>>>(pp);
    GRMP_CUDA(cudaGetLastError());
  }
  GRMP_CUDA(cudaFuncSetAttribute(p2tet_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(max_smem, 1024)));
  GRMP_CUDA(cudaStreamSynchronize(s));
  return GRMP_OK;
}

int fast_p2tet_numeric(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, const FastP2Tet& f, double* nzval) {
  if (f.ntiles == 0) return GRMP_OK;
  TileParams tp{p.g, pat.colptr.p, f.col_pairbeg.p, f.pairs.p, f.tile_colbeg.p, f.tile_cellbeg.p, f.tile_cells.p, p.factor, nzval};
  p2tet_tile_kernel<<<f.ntiles, TPB, f.smem_bytes, ctx->stream>>>(tp);
  GRMP_CUDA(cudaGetLastError());
  return GRMP_OK;
}

}  // namespace grmp
