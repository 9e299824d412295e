// fastpath_p2tet.cu -- owner-computes numeric kernels for the metric configuration:
// 3D P2 Laplace stiffness (H1P2{1,3}, [Gradient, Gradient], NoAction) on a frozen pattern.
//
// Replaces, for this form, the whole cell loop of assemble! (bilinearform.jl:226-377:
// update_trafo!/mapderiv!, update_basis! Gradient, quadrature contraction, _addnz scatter).
// Every stored non-zero is computed once, in registers, and written once:
//
//   edge kernel (p2tet_edge_kernel): one thread owns one EDGE column (pq) and walks the cells
//     around the edge in ring order.  With the exact P2 integrals every local column
//     K_loc[:, e_pq] is a short expression in S_ab = kappa |T| grad(lambda_a).grad(lambda_b)
//     (the order-2 rule of quadrature.jl:274-284 integrates them exactly).  Rows v_p, v_q, e_pq
//     receive a contribution from every ring cell and are summed in registers; the rows of a
//     ring vertex c (v_c, e_pc, e_qc) receive exactly two contributions from consecutive ring
//     cells and are completed through a 3-register carry; e_cc' has a single contribution.
//     Nothing is read-modify-written in memory.  The column is staged in shared memory and
//     written back with coalesced stores.  S of a tile's distinct cells is computed once per
//     tile into shared memory (the affine pullback of feevaluator_h1.jl:61-74 reduced to its
//     10 invariants).
//   VERTEX columns need no work of their own: the matrix of this form is symmetric, so every
//     off-diagonal entry (i, v_a) is the mirror image of an entry (v_a, i) that an edge thread
//     has in a register anyway (i an edge dof), or the ring sum -0.2*sum S_pq of the edge (a b)
//     (i = v_b).  Edge threads store those values straight to their mirrored slots.
//   diagonal kernel (p2tet_vertex_diag_kernel): rows of a stiffness matrix sum to zero, so
//     A[v,v] = -sum_{i != v} A[i,v]; one warp per vertex column, fixed reduction tree.
//
// No atomics, fixed summation orders -> deterministic.  Values agree with the reference's order of
// operations to rounding (tests/test_gpu_parity.py states the tolerance); the PATTERN always
// comes from the bit-exact symbolic pass, and slots are looked up in it by (row, column).
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "fastpath.cuh"

namespace grmp {

namespace {

constexpr int TPB_DEFAULT = 256;         // threads (= edge columns) per tile
constexpr int SMEM_BUDGET_DEFAULT = 88 * 1024;   // nzval stage + S of the tile's distinct cells
constexpr u32 NONE = 0xffffffffu;

// local edge e -> (p,q), Tetrahedron3D edges [1 2],[1 3],[1 4],[2 3],[2 4],[3 4] (h1_p2.jl:231-236)
__host__ __device__ inline void edge_nodes(int e, int& p, int& q) {
  const int P[6] = {0, 0, 0, 1, 1, 2}, Q[6] = {1, 2, 3, 2, 3, 3};
  p = P[e]; q = Q[e];
}
__host__ __device__ inline int edge_of(int a, int b) {   // local edge dof index 4.. of the vertex pair
  if (a > b) { int t = a; a = b; b = t; }
  return 4 + ((a == 0) ? (b - 1) : (a == 1 ? b + 1 : 5));
}
__host__ __device__ inline int sidx(int a, int b) {      // index of S_ab in the packed upper triangle
  if (a > b) { int t = a; a = b; b = t; }
  return a * 4 - a * (a - 1) / 2 + (b - a);
}

// S_ab = factor * |T| * grad(lambda_a).grad(lambda_b), packed (00,01,02,03,11,12,13,22,23,33)
// one 256-bit gather per node from the padded coordinate copy (a node is 24 B inside one 32 B sector)
__device__ __forceinline__ void load_node(const double* coords4, int node1, double& x, double& y, double& z) {
  double w;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(coords4 + 4 * (i64)(node1 - 1)));
}
__device__ __forceinline__ void cell_S(const GridView& g, const int4 nd, double factor, double* S, int dbg = 0) {
  double p0x, p0y, p0z, p1x, p1y, p1z, p2x, p2y, p2z, p3x, p3y, p3z;
  if (!(dbg & 4)) {   // default: three 64-bit gathers per node (measured 8 % faster than one 256-bit gather from a padded copy)
    const double* x0 = g.coords + (i64)(nd.x - 1) * 3; const double* x1 = g.coords + (i64)(nd.y - 1) * 3;
    const double* x2 = g.coords + (i64)(nd.z - 1) * 3; const double* x3 = g.coords + (i64)(nd.w - 1) * 3;
    p0x = x0[0]; p0y = x0[1]; p0z = x0[2]; p1x = x1[0]; p1y = x1[1]; p1z = x1[2];
    p2x = x2[0]; p2y = x2[1]; p2z = x2[2]; p3x = x3[0]; p3y = x3[1]; p3z = x3[2];
  } else {
  load_node(g.coords4, nd.x, p0x, p0y, p0z);
  load_node(g.coords4, nd.y, p1x, p1y, p1z);
  load_node(g.coords4, nd.z, p2x, p2y, p2z);
  load_node(g.coords4, nd.w, p3x, p3y, p3z);
  }
  const double ax = p1x - p0x, ay = p1y - p0y, az = p1z - p0z;
  const double bx = p2x - p0x, by = p2y - p0y, bz = p2z - p0z;
  const double cx = p3x - p0x, cy = p3y - p0y, cz = p3z - p0z;
  // n1 = b x c, n2 = c x a, n3 = a x b : grad(lambda_k) = n_k / det
  const double n1x = by * cz - bz * cy, n1y = bz * cx - bx * cz, n1z = bx * cy - by * cx;
  const double n2x = cy * az - cz * ay, n2y = cz * ax - cx * az, n2z = cx * ay - cy * ax;
  const double n3x = ay * bz - az * by, n3y = az * bx - ax * bz, n3z = ax * by - ay * bx;
  const double n0x = -(n1x + n2x + n3x), n0y = -(n1y + n2y + n3y), n0z = -(n1z + n2z + n3z);
  // |T| / det^2 = 1 / (6 |det|): the cell volume is recomputed from the coordinates (|det|/6), which agrees with
  // CellVolumes to rounding and saves a dependent load
  const double det = ax * n1x + ay * n1y + az * n1z;
  const double sc = factor / (6.0 * fabs(det));
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

// ---- records ---------------------------------------------------------------------------------
// pair record (16 B), pairs of a column stored in ring order:
//   x : tile-local cell | perm << 16 | flags << 24   (perm = p | q<<2 | in<<4 | out<<6 local vertex ids)
//   y : slot offsets inside the column of rows v_in, e_P,in, e_Q,in, e_in,out (255 = not in the pattern)
//   z : mirrored slot (global nzval index) of row e_PQ in column v_in, or NONE
//   w : chain-end pairs of multi-chain (halo) columns: mirrored slot of row e_PQ in column v_out, else unused
// column record (32 B):
//   a.x : offsets of rows v_P, v_Q, e_PQ | flags << 24 (bit 0: closed ring)
//   a.y : closing offsets (closed: rows of the first pair's in-vertex; open: rows of the last pair's out-vertex)
//   a.z, a.w : mirrored slots of row e_PQ in columns v_P, v_Q
//   b.x : mirrored slot of row e_PQ in the column of the closing vertex
//   b.y, b.z : slots of (row v_Q, col v_P) and (row v_P, col v_Q);  b.w : first pair;  a.y >> 24 : number of pairs
constexpr u32 PF_FIRST = 1u;   // first pair of the column (closed ring: its in-rows are completed at the end)
constexpr u32 PF_RESET = 2u;   // first pair of a further chain (halo columns of a partition): drop the carry
constexpr u32 PF_END = 4u;     // last pair of a chain that is not the last chain: mirror its out-vertex row now (slot in w)

struct PackParams {
  const u32* pair_cell;     // global cell of the pair (ring order)
  const u32* pair_local;    // tile-local cell
  const u32* pair_code;     // perm | flags << 8
  const i64* col_pairbeg;
  const i64* colptr;        // 1-based
  const i64* rowval;        // 1-based
  const i32* celldofs;
  const u32* col_of_pair;   // column of every pair
  const unsigned char* col_closed;
  i64 npairs, ncols;
  uint4* pairs;
  uint4* cols;              // 2 per column
};

// slot of (row, col) in the pattern as an offset inside the column, or -1
__device__ __forceinline__ i64 find_slot(const PackParams& p, i64 row0, i64 col0) {
  i64 lo = p.colptr[col0] - 1, hi = p.colptr[col0 + 1] - 1;
  const i64 beg = lo, end = hi, target = row0 + 1;
  while (lo < hi) {
    i64 mid = (lo + hi) >> 1;
    if (p.rowval[mid] < target) lo = mid + 1; else hi = mid;
  }
  if (lo < end && p.rowval[lo] == target) return lo - beg;
  return -1;
}
__device__ __forceinline__ u32 off8(i64 o) { return (o < 0 || o > 254) ? 255u : (u32)o; }
__device__ __forceinline__ u32 gslot(const PackParams& p, i64 row0, i64 col0) {
  i64 o = find_slot(p, row0, col0);
  return o < 0 ? NONE : (u32)(p.colptr[col0] - 1 + o);
}

__global__ void pack_pairs(PackParams p) {
  i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (k >= p.npairs) return;
  const i64 col = p.col_of_pair[k];
  if (p.col_closed[col] == 2) { p.pairs[k] = make_uint4(0, 0, 0, 0); return; }   // vertex column: no pair work
  const i64 cell = p.pair_cell[k];
  const u32 code = p.pair_code[k];
  const int P = code & 3, Q = (code >> 2) & 3, I = (code >> 4) & 3, O = (code >> 6) & 3;
  const i32* d = p.celldofs + cell * 10;
  const i64 vin = d[I] - 1;
  uint4 rec;
  rec.x = p.pair_local[k] | ((code & 255u) << 16) | (((code >> 8) & 255u) << 24);
  rec.y = off8(find_slot(p, vin, col)) | (off8(find_slot(p, d[edge_of(P, I)] - 1, col)) << 8) |
          (off8(find_slot(p, d[edge_of(Q, I)] - 1, col)) << 16) | (off8(find_slot(p, d[edge_of(I, O)] - 1, col)) << 24);
  rec.z = gslot(p, col, vin);
  rec.w = (((code >> 8) & PF_END) != 0) ? gslot(p, col, d[O] - 1) : NONE;
  p.pairs[k] = rec;
}

__global__ void pack_cols(PackParams p) {
  i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (j >= p.ncols) return;
  const i64 kb = p.col_pairbeg[j], ke = p.col_pairbeg[j + 1];
  uint4 a = make_uint4(0x00ffffffu, 0x00ffffffu, NONE, NONE), b = make_uint4(NONE, NONE, NONE, 0);
  if (ke > kb && p.col_closed[j] != 2) {     // 2 = not an edge column
    const bool closed = p.col_closed[j] == 1;
    // reference orientation (P,Q) = first ring pair
    const u32 c0 = p.pair_code[kb];
    const i32* d0 = p.celldofs + (i64)p.pair_cell[kb] * 10;
    const i64 vP = d0[c0 & 3] - 1, vQ = d0[(c0 >> 2) & 3] - 1;
    a.x = off8(find_slot(p, vP, j)) | (off8(find_slot(p, vQ, j)) << 8) | (off8(find_slot(p, j, j)) << 16) | ((closed ? 1u : 0u) << 24);
    // closing vertex: closed -> in-vertex of the first pair; open -> out-vertex of the last pair
    const i64 kc = closed ? kb : ke - 1;
    const u32 cc = p.pair_code[kc];
    const i32* dc = p.celldofs + (i64)p.pair_cell[kc] * 10;
    const int Pc = cc & 3, Qc = (cc >> 2) & 3, Vc = closed ? ((cc >> 4) & 3) : ((cc >> 6) & 3);
    const i64 vC = dc[Vc] - 1;
    a.y = off8(find_slot(p, vC, j)) | (off8(find_slot(p, dc[edge_of(Pc, Vc)] - 1, j)) << 8) | (off8(find_slot(p, dc[edge_of(Qc, Vc)] - 1, j)) << 16) | ((u32)(ke - kb) << 24);
    a.z = gslot(p, j, vP);
    a.w = gslot(p, j, vQ);
    b.x = gslot(p, j, vC);
    b.y = gslot(p, vQ, vP);
    b.z = gslot(p, vP, vQ);
    b.w = (u32)kb;
  }
  p.cols[2 * j] = a;
  p.cols[2 * j + 1] = b;
}

struct EdgeParams {
  GridView g;
  const i64* colptr;        // 1-based [ncols+1]
  const i64* col_pairbeg;   // [ncols+1]
  const uint4* pairs;
  const uint4* cols;
  const uint2* spokes;      // per column: scratch slots of the two end vertices' spoke lists
  double* dscratch;         // [sum of spoke counts] 0.2 * ring sum of S_vv per (vertex, spoke)
  const int4* tile_hdr;     // 2 per tile: {first column, #columns, first tile cell, #tile cells}, {g0 lo, g0 hi, nnz, 0}
  const int4* tile_nodes;   // CellNodes of the tiles' distinct cells
  double factor;
  double* nzval;
  int dbg;                  // GRMP_DEBUG_FLAGS: bit 0 = skip the mirrored stores (timing experiments only)
};


template <int TPB>
__global__ void __launch_bounds__(TPB, (TPB <= 64 ? 10 : TPB <= 128 ? 5 : TPB <= 192 ? 3 : TPB <= 256 ? 2 : 1)) p2tet_edge_kernel(const EdgeParams p) {
  extern __shared__ double sm[];
  __shared__ uint4 s_tab[256];   // perm code -> byte offsets (k * nct * 8, 16 bit each) of S_pp,S_qq,S_pq,S_pi,S_po,S_qi,S_qo
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int4 h0 = p.tile_hdr[2 * tile], h1 = p.tile_hdr[2 * tile + 1];
  const int c0 = h0.x, ncol = h0.y, cb = h0.z, nct = h0.w;
  const i64 g0 = (i64)(u32)h1.x | ((i64)h1.y << 32);
  const int nnz_t = h1.z;
  // stage[i] mirrors nzval[g0 + i]; it is shifted by one element when g0 is odd so that shared and global addresses
  // of the same element are 16-byte aligned together (TMA bulk store).  Every slot is written exactly once -> no zero-init.
  const int odd = (int)(g0 & 1);
  double* __restrict__ stage = sm + odd;
  double* __restrict__ S = sm + nnz_t + 2;
  // node ids of this thread's tile cells (up to GC per thread), issued first: the coordinate gathers depend on them
  constexpr int GC = 4;
  int4 nd[GC];
#pragma unroll
  for (int r = 0; r < GC; r++) {
    const int i = tid + r * TPB;
    nd[r] = (i < nct) ? p.tile_nodes[cb + i] : make_int4(1, 1, 1, 1);
  }
  // this thread's column record, pair range and first batch of pair records
  const int col = c0 + tid;
  const bool has_col = tid < ncol;
  u32 kb = 0, ke = 0;
  int abase = 0;
  uint4 ca = make_uint4(0, 0, 0, 0), cbx = make_uint4(0, 0, 0, 0);
  uint2 spk = make_uint2(NONE, NONE);
  if (has_col) {
    ca = p.cols[2 * (i64)col]; cbx = p.cols[2 * (i64)col + 1];
    spk = p.spokes[col];
    abase = (int)(p.colptr[col] - 1 - g0);
    kb = cbx.w; ke = kb + (ca.y >> 24);
  }
  for (int c = tid; c < 256; c += TPB) {
    const u32 P = c & 3, Q = (c >> 2) & 3, I = (c >> 4) & 3, O = (c >> 6) & 3;
    const u32 m = (u32)nct * 8u;
    uint4 t;
    t.x = (sidx(P, P) * m) | ((sidx(Q, Q) * m) << 16);
    t.y = (sidx(P, Q) * m) | ((sidx(P, I) * m) << 16);
    t.z = (sidx(P, O) * m) | ((sidx(Q, I) * m) << 16);
    t.w = (sidx(Q, O) * m);
    s_tab[c] = t;
  }
  uint4 rnext = (kb < ke) ? p.pairs[kb] : make_uint4(0, 0, 0, 0);   // a column's records share one or two cache lines
  // ---- geometry of the tile's distinct cells ----
#pragma unroll
  for (int r = 0; r < GC; r++) {
    const int i = tid + r * TPB;
    if (i < nct) {
      double sA[10];
      cell_S(p.g, nd[r], p.factor, sA, p.dbg);
#pragma unroll
      for (int k = 0; k < 10; k++) S[k * nct + i] = sA[k];        // k-major: conflict-free stores
    }
  }
  for (int i = tid + GC * TPB; i < nct; i += TPB) {               // tiles with more than GC*TPB cells (rare)
    double sA[10];
    cell_S(p.g, p.tile_nodes[cb + i], p.factor, sA, p.dbg);
#pragma unroll
    for (int k = 0; k < 10; k++) S[k * nct + i] = sA[k];
  }
  __syncthreads();
  if (has_col && ke > kb) {
    double* __restrict__ a = stage + abase;
    double A = 0.0, B = 0.0, C = 0.0, W = 0.0;       // rows v_P, v_Q, e_PQ of the column; (v_P, v_Q) coupling
    double c0r = 0.0, c1r = 0.0, c2r = 0.0;          // carry: partial rows of the shared ring vertex
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;             // closed ring: in-rows of the first pair, completed at the end
    const bool closed = (ca.x >> 24) & 1u;
    const char* __restrict__ Sb = reinterpret_cast<const char*>(S);
    double Tp = 0.0, Tq = 0.0;                       // ring sums of S_PP, S_QQ: 0.2*sum over the spokes = diagonal of v_P, v_Q
#pragma unroll 2
    for (u32 k = kb; k < ke; k++) {
      const uint4 r = rnext;
      if (k + 1 < ke) rnext = p.pairs[k + 1];
      const uint4 t = s_tab[(r.x >> 16) & 255u];
      const char* sc = Sb + (r.x & 0xffffu) * 8u;
      const double spp = *reinterpret_cast<const double*>(sc + (t.x & 0xffffu));
      const double sqq = *reinterpret_cast<const double*>(sc + (t.x >> 16));
      const double spq = *reinterpret_cast<const double*>(sc + (t.y & 0xffffu));
      const double spi = *reinterpret_cast<const double*>(sc + (t.y >> 16));
      const double sqi = *reinterpret_cast<const double*>(sc + (t.z >> 16));
      // rows of S sum to zero (the barycentric gradients do): two of the seven values follow from the other five
      double spo, sqo;
      if (p.dbg & 8) {
        spo = *reinterpret_cast<const double*>(sc + (t.z & 0xffffu));
        sqo = *reinterpret_cast<const double*>(sc + (t.w & 0xffffu));
      } else {
        spo = -((spp + spq) + spi);
        sqo = -((sqq + spq) + sqi);
      }
      A += 0.6 * spq - 0.2 * spp;
      B += 0.6 * spq - 0.2 * sqq;
      C += 1.6 * (spp + sqq + spq);
      W += -0.2 * spq;
      Tp += spp; Tq += sqq;
      const double base = spq + spp, baseq = spq + sqq;
      const u32 fl = r.x >> 24;
      if (fl & PF_RESET) { c0r = 0.0; c1r = 0.0; c2r = 0.0; }
      const double in0 = c0r + -0.2 * (spi + sqi);                 // v_in
      const double in1 = c1r + 0.8 * (2.0 * sqi + spi + base);      // e_P,in
      const double in2 = c2r + 0.8 * (2.0 * spi + sqi + baseq);     // e_Q,in
      c0r = -0.2 * (spo + sqo);                                     // v_out
      c1r = 0.8 * (2.0 * sqo + spo + base);                         // e_P,out
      c2r = 0.8 * (2.0 * spo + sqo + baseq);                        // e_Q,out
      const double x = 0.8 * (spi + spo + sqi + sqo);               // e_in,out
      const u32 o3 = r.y >> 24;
      if (o3 != 255u) a[o3] = x;
      if ((fl & PF_END) && r.w != NONE && !(p.dbg & 1)) p.nzval[r.w] = c0r;         // chain end inside a halo column: mirror (e_PQ, v_out)
      if (closed && (fl & PF_FIRST)) {
        f0 = in0; f1 = in1; f2 = in2;                               // partner is the last pair of the ring
      } else {
        const u32 o0 = r.y & 255u, o1 = (r.y >> 8) & 255u, o2 = (r.y >> 16) & 255u;
        if (o0 != 255u) a[o0] = in0;
        if (o1 != 255u) a[o1] = in1;
        if (o2 != 255u) a[o2] = in2;
        if (r.z != NONE && !(p.dbg & 1)) p.nzval[r.z] = in0;                        // mirror (e_PQ, v_in)
      }
    }
    // closing rows: closed ring -> first pair's in-rows + last carry; open chain -> last pair's out-rows
    {
      const double q0 = closed ? f0 + c0r : c0r, q1 = closed ? f1 + c1r : c1r, q2 = closed ? f2 + c2r : c2r;
      const u32 o0 = ca.y & 255u, o1 = (ca.y >> 8) & 255u, o2 = (ca.y >> 16) & 255u;
      if (o0 != 255u) a[o0] = q0;
      if (o1 != 255u) a[o1] = q1;
      if (o2 != 255u) a[o2] = q2;
      if (cbx.x != NONE && !(p.dbg & 1)) p.nzval[cbx.x] = q0;
    }
    {
      const u32 oA = ca.x & 255u, oB = (ca.x >> 8) & 255u, oC = (ca.x >> 16) & 255u;
      if (oA != 255u) a[oA] = A;
      if (oB != 255u) a[oB] = B;
      if (oC != 255u) a[oC] = C;
      if (!(p.dbg & 1)) {
      if (ca.z != NONE) p.nzval[ca.z] = A;       // (e_PQ, v_P)
      if (ca.w != NONE) p.nzval[ca.w] = B;       // (e_PQ, v_Q)
      if (cbx.y != NONE) p.nzval[cbx.y] = W;     // (v_Q, v_P)
      if (cbx.z != NONE) p.nzval[cbx.z] = W;     // (v_P, v_Q)
      }
      if (spk.x != NONE) p.dscratch[spk.x] = 0.2 * Tp;
      if (spk.y != NONE) p.dscratch[spk.y] = 0.2 * Tq;
    }
  }
  __syncthreads();
  {
    // the tile's nzval range is contiguous: one TMA bulk store (cp.async.bulk shared -> global) of the 16-byte aligned
    // body, the (at most one) unaligned element at either end by ordinary stores
    double* __restrict__ dst = p.nzval + g0;
    const int i0 = odd;                                   // first element whose address is 16-byte aligned
    const int nb = (nnz_t > i0) ? ((nnz_t - i0) & ~1) : 0;  // elements in the bulk body
    if (tid == 0 && nb > 0) {
      const unsigned src = (unsigned)__cvta_generic_to_shared(stage + i0);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + i0), "r"(src), "r"(nb * 8) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (tid == 1 && i0 == 1 && nnz_t > 0) dst[0] = stage[0];
    if (tid == 2 && i0 + nb < nnz_t) dst[i0 + nb] = stage[i0 + nb];
    if (tid == 0 && nb > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must stay alive until read
  }
}

// A[v,v] = 0.6 * sum_{K containing v} S_vv = 0.2 * sum over the spokes (v w) of the ring sums of S_vv (every cell at v
// has three edges at v); the spoke values were left in dscratch by the edge threads.  Fixed order -> deterministic.
__global__ void p2tet_vertex_diag_kernel(const uint4* vrec, i64 nv, const double* __restrict__ dscratch, double* nzval) {
  const i64 w = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (w >= nv) return;
  const uint4 r = vrec[w];                 // {diagonal slot | NONE, first spoke slot, #spokes, 0}
  if (r.x == NONE) return;
  double s = 0.0;
  for (u32 k = 0; k < r.z; k++) s += dscratch[r.y + k];
  nzval[r.x] = s;
}

__global__ void find_diag_slots(const u32* vcols, const u32* vspoke, i64 nv, const i64* colptr, const i64* rowval, uint4* vrec) {
  i64 w = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (w >= nv) return;
  const i64 col = vcols[w];
  i64 lo = colptr[col] - 1, hi = colptr[col + 1] - 1;
  const i64 end = hi;
  while (lo < hi) {
    i64 mid = (lo + hi) >> 1;
    if (rowval[mid] < col + 1) lo = mid + 1; else hi = mid;
  }
  const u32 d = (lo < end && rowval[lo] == col + 1) ? (u32)lo : NONE;
  vrec[w] = make_uint4(d, vspoke[2 * w], vspoke[2 * w + 1], 0);
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
         p.e1.tab_nd == 10 && p.action == GRMP_ACT_NONE && p.reg.n == 0 &&
         (p.apt == GRMP_APT_SYMMETRIC || (p.apt == GRMP_APT_BILINEARFORM && !p.transposed));
}

int fast_p2tet_build(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, const std::vector<double>& w,
                     const std::vector<double>& derivs, i64 ncols_owned, FastP2Tet* out) {
  // halo columns (>= ncols_owned) are processed too: their mirrors complete the owned vertex columns;
  // only they may consist of several chains (cells around a halo edge are present only where they touch an owned dof)
  cudaStream_t s = ctx->stream;
  const i64 ncells = p.g.ncells, ncols = pat.ncols;
  const i64 ncols_owned_eff = (ncols_owned >= 0 && ncols_owned < ncols) ? ncols_owned : ncols;
  out->ntiles = 0; out->nvcols = 0;
  if (pat.nnz >= (i64)NONE) return fail(GRMP_EUNSUPPORTED, "fast path: more than 2^32-1 non-zeros on one device");
  // (0) the caller's tables must be the standard P2 basis integrated exactly
  {
    const int nq = p.nq;
    if ((int)w.size() != nq || derivs.size() != (size_t)nq * 3 * 10) return fail(GRMP_EUNSUPPORTED, "fast path: table shape");
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
  std::vector<i32> h_cn((size_t)ncells * 4), h_dofs((size_t)ncells * 10);
  GRMP_CUDA(cudaMemcpyAsync(h_cell.data(), dg.gcell.p, npairs * 4, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaMemcpyAsync(h_src.data(), dg.gsrc.p, npairs * 4, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaMemcpyAsync(h_pairbeg.data(), dg.segptr.p, (ncols + 1) * 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaMemcpyAsync(h_colptr.data(), pat.colptr.p, (ncols + 1) * 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaMemcpyAsync(h_cn.data(), p.g.cellnodes, (size_t)ncells * 16, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaMemcpyAsync(h_dofs.data(), p.e1.celldofs, (size_t)ncells * 40, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  // tile shape (tunable for experiments: GRMP_FAST_TPB in {64,128,256}, GRMP_FAST_SMEM_KB)
  int TPB = getenv("GRMP_FAST_TPB") ? atoi(getenv("GRMP_FAST_TPB")) : TPB_DEFAULT;
  if (TPB != 64 && TPB != 128 && TPB != 192 && TPB != 256 && TPB != 512) TPB = TPB_DEFAULT;
  const i64 SMEM_BUDGET = getenv("GRMP_FAST_SMEM_KB") ? 1024 * (i64)atoi(getenv("GRMP_FAST_SMEM_KB")) : SMEM_BUDGET_DEFAULT * (i64)TPB / TPB_DEFAULT;
  out->tpb = TPB;
  // (2) host: ring order of every edge column, tiles over the edge columns, list of vertex columns
  std::vector<u32> pair_cell(npairs), pair_local(npairs), pair_code(npairs), col_of_pair(npairs), vcols;
  std::vector<unsigned char> col_closed(ncols, 2);
  std::vector<u32> endP(ncols, NONE), endQ(ncols, NONE);     // vertex dofs of the two ends of every edge column
  std::vector<int4> tile_hdr, tile_nodes;
  std::vector<i32> tile_cells;   // distinct cells of the open tile (global ids), flushed into tile_nodes
  std::vector<i32> mark(ncells, -1), local_of(ncells, 0);
  tile_nodes.reserve((size_t)ncells * 3);
  int cur_tile = 0, cur_cols = 0, cur_cells = 0;
  i64 cur_nnz = 0, tile_first_col = 0;
  i64 max_smem = 0;
  auto close_tile = [&](i64 end_col) {
    if (cur_cols == 0) return;
    const i64 g0 = h_colptr[tile_first_col] - 1;
    tile_hdr.push_back(make_int4((int)tile_first_col, (int)(end_col - tile_first_col), (int)tile_nodes.size(), cur_cells));
    tile_hdr.push_back(make_int4((int)(u32)(g0 & 0xffffffffll), (int)(g0 >> 32), (int)cur_nnz, 0));
    for (i32 c : tile_cells) tile_nodes.push_back(make_int4(h_cn[(size_t)c * 4], h_cn[(size_t)c * 4 + 1], h_cn[(size_t)c * 4 + 2], h_cn[(size_t)c * 4 + 3]));
    tile_cells.clear();
    max_smem = std::max<i64>(max_smem, 8 * cur_nnz + 80 * (i64)cur_cells + 32);
    cur_tile++; cur_cols = 0; cur_cells = 0; cur_nnz = 0;
  };
  struct RP { u32 cell; int P, Q, R, S; i32 nR, nS; };
  std::vector<RP> rp;
  std::vector<char> used;
  for (i64 j = 0; j < ncols; j++) {
    const i64 kb = h_pairbeg[j], ke = h_pairbeg[j + 1];
    const i64 len = h_colptr[j + 1] - h_colptr[j];
    if (ke == kb) { close_tile(j); continue; }
    const int lj0 = (int)(h_src[kb] / (u32)ncells);
    if (lj0 < 4) {   // vertex column: filled by mirrors + the diagonal kernel
      close_tile(j);
      vcols.push_back((u32)j);
      for (i64 k = kb; k < ke; k++) { pair_cell[k] = h_cell[k]; pair_local[k] = 0; pair_code[k] = 0; col_of_pair[k] = (u32)j; }
      continue;
    }
    if (len > 254 || ke - kb > 255) return fail(GRMP_EUNSUPPORTED, "fast path: an edge column has more than 254 entries");
    // ---- ring order of the cells around the edge ----
    const int n = (int)(ke - kb);
    rp.resize(n);
    i32 P0 = 0;
    for (int t = 0; t < n; t++) {
      const u32 c = h_cell[kb + t];
      const int lj = (int)(h_src[kb + t] / (u32)ncells);
      if (lj < 4) return fail(GRMP_EUNSUPPORTED, "fast path: mixed dof types in one column");
      int pl, ql; edge_nodes(lj - 4, pl, ql);
      int rl = -1, sl = -1;
      for (int v = 0; v < 4; v++) if (v != pl && v != ql) { if (rl < 0) rl = v; else sl = v; }
      const i32* cn = &h_cn[(size_t)c * 4];
      if (t == 0) P0 = cn[pl];
      if (cn[pl] != P0) { int tmp = pl; pl = ql; ql = tmp; }      // consistent global orientation (P,Q)
      if (cn[pl] != P0) return fail(GRMP_EUNSUPPORTED, "fast path: inconsistent edge column");
      rp[t] = RP{c, pl, ql, rl, sl, cn[rl], cn[sl]};
      if (t == 0) { endP[j] = (u32)(h_dofs[(size_t)c * 10 + pl] - 1); endQ[j] = (u32)(h_dofs[(size_t)c * 10 + ql] - 1); }
    }
    // degrees of the ring vertices; chains start at vertices of degree 1, a star without such a vertex is a closed ring
    std::vector<int> degR(n), degS(n);
    bool closed = true;
    for (int t = 0; t < n; t++) {
      int dR = 0, dS = 0;
      for (int u = 0; u < n; u++) {
        dR += (rp[u].nR == rp[t].nR) + (rp[u].nS == rp[t].nR);
        dS += (rp[u].nR == rp[t].nS) + (rp[u].nS == rp[t].nS);
      }
      if (dR > 2 || dS > 2) return fail(GRMP_EUNSUPPORTED, "fast path: non-manifold edge star");
      degR[t] = dR; degS[t] = dS;
      if (dR == 1 || dS == 1) closed = false;
    }
    used.assign(n, 0);
    int step = 0, nchains = 0;
    while (step < n) {
      int start = -1, start_in_is_R = 1;
      if (closed) { if (step != 0) return fail(GRMP_EUNSUPPORTED, "fast path: edge star is not a single ring"); start = 0; }
      else
        for (int t = 0; t < n && start < 0; t++)
          if (!used[t]) { if (degR[t] == 1) { start = t; start_in_is_R = 1; } else if (degS[t] == 1) { start = t; start_in_is_R = 0; } }
      if (start < 0) return fail(GRMP_EUNSUPPORTED, "fast path: edge star mixes a ring and chains");
      if (nchains > 0 && j < ncols_owned_eff) return fail(GRMP_EUNSUPPORTED, "fast path: an owned edge star is not a single chain");
      int curp = start;
      const i32 first_in = start_in_is_R ? rp[start].nR : rp[start].nS;
      i32 vin = first_in;
      bool chain_first = true;
      while (true) {
        used[curp] = 1;
        const bool inR = (rp[curp].nR == vin);
        const int I = inR ? rp[curp].R : rp[curp].S, O = inR ? rp[curp].S : rp[curp].R;
        const i32 vout = inR ? rp[curp].nS : rp[curp].nR;
        const i64 k = kb + step;
        u32 fl = (step == 0) ? PF_FIRST : 0u;
        if (chain_first && nchains > 0) fl |= PF_RESET;
        pair_cell[k] = rp[curp].cell;
        pair_code[k] = (u32)(rp[curp].P | (rp[curp].Q << 2) | (I << 4) | (O << 6)) | (fl << 8);
        col_of_pair[k] = (u32)j;
        step++; chain_first = false;
        int nxt = -1;
        for (int u = 0; u < n; u++) if (!used[u] && (rp[u].nR == vout || rp[u].nS == vout)) { nxt = u; break; }
        if (nxt < 0) {
          if (closed && (step != n || vout != first_in)) return fail(GRMP_EUNSUPPORTED, "fast path: edge ring does not close");
          if (!closed && step < n) pair_code[k] |= (PF_END << 8);     // a further chain follows
          break;
        }
        vin = vout; curp = nxt;
      }
      nchains++;
    }
    col_closed[j] = closed ? 1 : 0;
    // ---- tile budget ----
    for (int attempt = 0; attempt < 2; attempt++) {
      int fresh = 0;
      for (i64 k = kb; k < ke; k++) if (mark[pair_cell[k]] != cur_tile) fresh++;
      const i64 need = 8 * (cur_nnz + len) + 80 * (i64)(cur_cells + fresh);
      if (cur_cols > 0 && (cur_cols + 1 > TPB || need > SMEM_BUDGET)) { close_tile(j); continue; }
      if (cur_cols == 0) tile_first_col = j;
      for (i64 k = kb; k < ke; k++) {
        const u32 c = pair_cell[k];
        if (mark[c] != cur_tile) { mark[c] = cur_tile; local_of[c] = cur_cells++; tile_cells.push_back((i32)c); }
        pair_local[k] = (u32)local_of[c];
      }
      cur_cols++; cur_nnz += len;
      break;
    }
  }
  close_tile(ncols);
  if (max_smem > 220 * 1024) return fail(GRMP_EUNSUPPORTED, "fast path: a single column exceeds the shared-memory tile");
  const int ntiles = (int)(tile_hdr.size() / 2);
  if (npairs >= (i64)NONE) return fail(GRMP_EUNSUPPORTED, "fast path: more than 2^32-1 pairs");
  out->ntiles = ntiles; out->npairs = npairs; out->smem_bytes = (int)max_smem; out->nvcols = (i64)vcols.size();
  if (tile_hdr.empty()) tile_hdr.assign(2, make_int4(0, 0, 0, 0));
  if (tile_nodes.empty()) tile_nodes.push_back(make_int4(1, 1, 1, 1));
  if (vcols.empty()) vcols.push_back(0);
  GRMP_TRY(out->tile_hdr.upload(tile_hdr.data(), tile_hdr.size(), s));
  GRMP_TRY(out->tile_nodes.upload(tile_nodes.data(), tile_nodes.size(), s));
  // (3) pack pair / column records on the device (slots are looked up in the pattern by (row, col))
  DevBuf<u32> d_cell, d_local, d_code, d_colof;
  DevBuf<unsigned char> d_closed;
  GRMP_TRY(d_cell.upload(pair_cell.data(), npairs, s)); GRMP_TRY(d_local.upload(pair_local.data(), npairs, s));
  GRMP_TRY(d_code.upload(pair_code.data(), npairs, s)); GRMP_TRY(d_colof.upload(col_of_pair.data(), npairs, s));
  GRMP_TRY(d_closed.upload(col_closed.data(), ncols, s));
  GRMP_TRY(out->pairs.alloc(std::max<i64>(npairs, 1)));
  GRMP_TRY(out->cols.alloc(2 * (size_t)std::max<i64>(ncols, 1)));
  GRMP_TRY(out->col_pairbeg.alloc(ncols + 1));
  GRMP_CUDA(cudaMemcpyAsync(out->col_pairbeg.p, dg.segptr.p, (ncols + 1) * 8, cudaMemcpyDeviceToDevice, s));
  PackParams pp{d_cell.p, d_local.p, d_code.p, dg.segptr.p, pat.colptr.p, pat.rowval.p, p.e1.celldofs, d_colof.p, d_closed.p,
                npairs, ncols, out->pairs.p, out->cols.p};
  if (npairs) pack_pairs<<<(unsigned)((npairs + 255) / 256), 256, 0, s>>>(pp);
  if (ncols) pack_cols<<<(unsigned)((ncols + 255) / 256), 256, 0, s>>>(pp);
  GRMP_CUDA(cudaGetLastError());
  // (3b) spoke lists: the diagonal of a vertex column is 0.2 * sum over its spokes of the ring sums of S_vv
  std::vector<u32> spoke_cnt(ncols, 0), spoke_ptr(ncols + 1, 0);
  for (i64 j = 0; j < ncols; j++) if (endP[j] != NONE) { spoke_cnt[endP[j]]++; spoke_cnt[endQ[j]]++; }
  for (i64 j = 0; j < ncols; j++) spoke_ptr[j + 1] = spoke_ptr[j] + spoke_cnt[j];
  std::vector<uint2> spokes(std::max<i64>(ncols, 1), make_uint2(NONE, NONE));
  std::fill(spoke_cnt.begin(), spoke_cnt.end(), 0);
  for (i64 j = 0; j < ncols; j++) if (endP[j] != NONE) {
    spokes[j].x = spoke_ptr[endP[j]] + spoke_cnt[endP[j]]++;
    spokes[j].y = spoke_ptr[endQ[j]] + spoke_cnt[endQ[j]]++;
  }
  std::vector<u32> vspoke(2 * vcols.size());
  for (size_t w2 = 0; w2 < vcols.size(); w2++) { vspoke[2 * w2] = spoke_ptr[vcols[w2]]; vspoke[2 * w2 + 1] = spoke_cnt[vcols[w2]]; }
  DevBuf<u32> d_vspoke;
  GRMP_TRY(d_vspoke.upload(vspoke.data(), vspoke.size(), s));
  GRMP_TRY(out->spokes.upload(spokes.data(), spokes.size(), s));
  GRMP_TRY(out->dscratch.alloc(std::max<size_t>(spoke_ptr[ncols], 1)));
  // (4) vertex columns: list + diagonal slots
  GRMP_TRY(out->vcols.upload(vcols.data(), vcols.size(), s));
  GRMP_TRY(out->vrec.alloc(vcols.size()));
  if (out->nvcols > 0) {
    find_diag_slots<<<(unsigned)((out->nvcols + 255) / 256), 256, 0, s>>>(out->vcols.p, d_vspoke.p, out->nvcols, pat.colptr.p, pat.rowval.p, out->vrec.p);
    GRMP_CUDA(cudaGetLastError());
  }
  const int smem_attr = (int)std::max<i64>(max_smem, 1024);
  GRMP_CUDA(cudaFuncSetAttribute(p2tet_edge_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr));
  GRMP_CUDA(cudaFuncSetAttribute(p2tet_edge_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr));
  GRMP_CUDA(cudaFuncSetAttribute(p2tet_edge_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr));
  GRMP_CUDA(cudaFuncSetAttribute(p2tet_edge_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr));
  GRMP_CUDA(cudaFuncSetAttribute(p2tet_edge_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr));
  GRMP_CUDA(cudaStreamSynchronize(s));
  return GRMP_OK;
}

int fast_p2tet_numeric(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, const FastP2Tet& f, double* nzval) {
  if (f.ntiles > 0) {
    static const int dbg = getenv("GRMP_DEBUG_FLAGS") ? atoi(getenv("GRMP_DEBUG_FLAGS")) : 0;
    EdgeParams ep{p.g, pat.colptr.p, f.col_pairbeg.p, f.pairs.p, f.cols.p, f.spokes.p, f.dscratch.p, f.tile_hdr.p, f.tile_nodes.p, p.factor, nzval, dbg};
    if (f.tpb == 64) p2tet_edge_kernel<64><<<f.ntiles, 64, f.smem_bytes, ctx->stream>>>(ep);
    else if (f.tpb == 256) p2tet_edge_kernel<256><<<f.ntiles, 256, f.smem_bytes, ctx->stream>>>(ep);
    else if (f.tpb == 192) p2tet_edge_kernel<192><<<f.ntiles, 192, f.smem_bytes, ctx->stream>>>(ep);
    else if (f.tpb == 512) p2tet_edge_kernel<512><<<f.ntiles, 512, f.smem_bytes, ctx->stream>>>(ep);
    else p2tet_edge_kernel<128><<<f.ntiles, 128, f.smem_bytes, ctx->stream>>>(ep);
    GRMP_CUDA(cudaGetLastError());
  }
  if (f.nvcols > 0 && !(getenv("GRMP_DEBUG_FLAGS") && (atoi(getenv("GRMP_DEBUG_FLAGS")) & 2))) {
    p2tet_vertex_diag_kernel<<<(unsigned)((f.nvcols + 255) / 256), 256, 0, ctx->stream>>>(f.vrec.p, f.nvcols, f.dscratch.p, nzval);
    GRMP_CUDA(cudaGetLastError());
  }
  return GRMP_OK;
}

}  // namespace grmp
