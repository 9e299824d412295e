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
//     Nothing is read-modify-written in memory.
//     Geometry lives in registers: the ring cell (p, q, c_k, c_k+1) shares p, q with the whole
//     ring and c_k with the previous cell, so ONE new vertex per cell is fetched (from the
//     tile's node coordinates in shared memory) and one cross product a x (c - p) is carried
//     (the affine pullback of feevaluator_h1.jl:61-74 reduced to the five off-diagonal
//     invariants S_pq, S_pi, S_po, S_qi, S_qo; the diagonal ones follow from the zero row sums).
//     A tile = up to NW groups of up to 32 consecutive edge columns, one group per consumer warp.
//     Its header, records, node coordinates and mirror list form one contiguous blob that a
//     single TMA bulk load (cp.async.bulk + mbarrier) brings into shared memory.  CTAs are
//     persistent and claim tiles from a global counter; one service warp keeps a 2-deep ring of
//     input blobs filled, the consumer warps never meet at a CTA-wide barrier.  Every warp
//     stages the contiguous nzval range of its group in its own shared-memory slot and writes
//     it with its own TMA bulk store, which overlaps the ring walk of its next group.
//   VERTEX columns need no ring walk of their own: the matrix of this form is symmetric, so every
//     off-diagonal entry (i, v_a) is the mirror image of an entry (v_a, i) that an edge thread
//     has in a register anyway (i an edge dof), or the ring sum -0.2*sum S_pq of the edge (a b)
//     (i = v_b).  Edge threads park those values in shared memory (in the slots of their already
//     consumed records); when all warps are done with a tile the service warp writes them to
//     their mirrored slots in DESTINATION order (list sorted in the symbolic pass), so that the
//     tile's values for one vertex column form one run of consecutive addresses.  Scattered
//     8-byte stores are the scarce resource of this kernel: the device retires ~1e11 partial
//     32-byte sector writes per second, half the rate of full sectors (tools/write_bw.cu).
//   diagonal kernel (p2tet_vertex_diag_kernel): the vertex-vertex block is the P1 stiffness matrix
//     scaled entrywise by 0.6 (diagonal) / -0.2 (off-diagonal) and P1 columns sum to zero, so
//     A[v,v] = 3 * sum of the vertex rows w != v of column v; eight lanes per column, fixed tree.
//
// No atomics on values, fixed summation orders -> deterministic.  Values agree with the reference's
// order of operations to rounding (tests/test_gpu_parity.py states the tolerance); the PATTERN
// always comes from the bit-exact symbolic pass, and slots are looked up in it by (row, column).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>

#include <cub/cub.cuh>

#include "fastpath.cuh"

namespace grmp {

namespace {

constexpr int NW_DEFAULT = 7;                     // consumer warps per CTA
constexpr int NSVC = 1;                           // service warps per CTA (tile loads, mirror write-out)
constexpr int SLOT_DEFAULT = 800;                 // nzval entries a warp may stage per group (32 columns x ~23 rows)
__host__ inline i64 smem_budget_default(int nw) { return 1024 * (i64)(nw >= 5 ? 112 : nw == 4 ? 74 : 55); }   // 2 / 3 / 4 CTAs per SM
constexpr u32 NONE = 0xffffffffu;
constexpr int MAX_TILE_NODES = 4095;              // 12-bit tile-local node ids

// local edge e -> (p,q), Tetrahedron3D edges [1 2],[1 3],[1 4],[2 3],[2 4],[3 4] (h1_p2.jl:231-236)
__host__ __device__ inline void edge_nodes(int e, int& p, int& q) {
  const int P[6] = {0, 0, 0, 1, 1, 2}, Q[6] = {1, 2, 3, 2, 3, 3};
  p = P[e]; q = Q[e];
}
__host__ __device__ inline int edge_of(int a, int b) {   // local edge dof index 4.. of the vertex pair
  if (a > b) { int t = a; a = b; b = t; }
  return 4 + ((a == 0) ? (b - 1) : (a == 1 ? b + 1 : 5));
}

// ---- records ---------------------------------------------------------------------------------
// tile blob (global, contiguous per tile, 16-byte aligned sections; one TMA bulk load):
//   tile header 48 B: see TileHdr;  group table (NW+1) x 16 B {first column (tile-local), first nzval slot (relative to g0),
//     first pair record (tile-local), -}
//   column records  32 B x ncol   (16 B used; after the ring walk the owner parks 4 doubles A, B, q0, W in its record)
//     x : tile-local node of P | Q << 12 | #pairs << 24
//     y : slot offsets inside the column of rows v_P, v_Q, e_PQ (255 = not in the pattern) | flags << 24 (bit 0: closed ring)
//     z : closing offsets (closed: rows of the first pair's in-vertex; open: rows of the last pair's out-vertex)
//     w : column start inside the group's nzval range | first pair (tile-local) << 16
//   pair records 8 B x npairs (the owner parks the value of row v_in in it).  Inside a group they are stored ROUND-major and
//     compact: first the records of ring position 0 of all columns of the group (lane order), then position 1 of the columns
//     that have one, ...  A warp reads / parks consecutive addresses in every round (no bank conflicts); a lane finds its
//     record with a ballot of "my ring is longer than k" and a running round base.
//     x : slot offsets inside the column of rows v_in, e_P,in, e_Q,in, e_in,out (255 = not in the pattern)
//     y : tile-local node of the in-vertex | out-vertex << 12 | flags << 24
//   node coordinates 24 B x nnodes : tile-blocked copy of Coordinates (the grid is immutable after grmp_grid_create)
//   mirror list, sorted by destination: u32 destination slot x nmax, then u16 source (8-byte word inside the blob) x nmax;
//     nmax = 5 ncol + npairs candidates, the first nmir (header) are valid.  Candidates of column c: (e_PQ, v_P) <- A,
//     (e_PQ, v_Q) <- B, (e_PQ, v_closing) <- q0, (v_Q, v_P) <- W, (v_P, v_Q) <- W; of pair k: (e_PQ, v_in) <- parked value.
constexpr u32 PF_RESET = 2u;   // first pair of a further chain (halo columns of a partition): drop the carry, reload the in-vertex
constexpr u32 PF_END = 4u;     // last pair of a chain that is not the last chain: mirror its out-vertex row now (slot in end_slots)

struct __align__(16) TileHdr {
  int c0, ncol, nnodes, npairs;          // first column, #columns, #distinct nodes, #pairs
  u32 g0lo, g0hi; int nnz; u32 blob16;   // first nzval slot, #slots, blob offset in 16-byte units
  u32 pair_base, node_base;              // global index of the first pair / node-list entry (blob copy: node_base holds nmir)
  u32 blob_bytes, cols_off;              // blob size; byte offset of the column records
};
static_assert(sizeof(TileHdr) == 48, "TileHdr layout");

__host__ __device__ inline u32 pad16(u32 b) { return (b + 15u) & ~15u; }
// section offsets inside a blob, all derived from the header
__host__ __device__ inline u32 off_pairs(u32 cols_off, u32 ncol) { return cols_off + 32u * ncol; }
__host__ __device__ inline u32 off_xyz(u32 cols_off, u32 ncol, u32 npairs) { return off_pairs(cols_off, ncol) + pad16(8u * npairs); }
__host__ __device__ inline u32 off_mdst(u32 cols_off, u32 ncol, u32 npairs, u32 nnodes) { return off_xyz(cols_off, ncol, npairs) + pad16(24u * nnodes); }
__host__ __device__ inline u32 off_msrc(u32 cols_off, u32 ncol, u32 npairs, u32 nnodes) { return off_mdst(cols_off, ncol, npairs, nnodes) + pad16(4u * (5u * ncol + npairs)); }
__host__ __device__ inline u32 blob_size(u32 cols_off, u32 ncol, u32 npairs, u32 nnodes) { return off_msrc(cols_off, ncol, npairs, nnodes) + pad16(2u * (5u * ncol + npairs)); }

struct PackParams {
  const u32* pair_cell;     // global cell of the pair (ring order)
  const u32* pair_io;       // tile-local in-node | out-node << 12
  const u32* pair_code;     // perm (P | Q<<2 | I<<4 | O<<6) | flags << 8
  const i64* col_pairbeg;
  const i64* colptr;        // 1-based
  const i64* rowval;        // 1-based
  const i32* celldofs;
  const u32* col_of_pair;   // column of every pair
  const unsigned char* col_closed;   // 0 open chain(s), 1 closed ring, 2 not an edge column
  const u32* col_tile;      // tile of every edge column
  const u32* col_pq;        // tile-local node of P | Q << 12
  const u32* col_abase;     // first slot of the column relative to its group
  const u32* col_group;     // first column of the column's group | #columns of the group << 27... see pack_pairs
  const u32* col_gcount;    // #columns of the column's group
  const TileHdr* hdr;
  const int* mir_base;      // [ntiles+1] first mirror candidate of every tile
  i64 npairs, ncols;
  unsigned char* blob;
  u32* end_slots;           // [npairs] or null
  u32 halo_first;           // mirrors into columns another rank owns (slots >= halo_first) are dropped
  u32* mkey;                // mirror candidates: destination slot (NONE = not in the pattern / no value)
  u32* mval;                //                    source word inside the blob
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
  if (o < 0) return NONE;
  const u32 g = (u32)(p.colptr[col0] - 1 + o);
  return g >= p.halo_first ? NONE : g;     // halo columns are not assembled on this rank
}

__global__ void pack_pairs(PackParams p) {
  i64 k = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (k >= p.npairs) return;
  const i64 col = p.col_of_pair[k];
  if (p.col_closed[col] == 2) return;   // vertex column: no pair work
  const i64 cell = p.pair_cell[k];
  const u32 code = p.pair_code[k];
  const int P = code & 3, Q = (code >> 2) & 3, I = (code >> 4) & 3, O = (code >> 6) & 3;
  const u32 fl = (code >> 8) & 255u;
  const i32* d = p.celldofs + cell * 10;
  const i64 vin = d[I] - 1;
  const u32 tile = p.col_tile[col];
  const TileHdr h = p.hdr[tile];
  // round-major compact position inside the group: rounds before this one + columns of the group in front with a ring this long
  u32 kl;
  {
    const i64 gf = p.col_group[col], gc = p.col_gcount[col];
    const i64 kr = k - p.col_pairbeg[col];
    u32 idx = 0;
    for (i64 c = gf; c < gf + gc; c++) {
      const i64 npc = p.col_pairbeg[c + 1] - p.col_pairbeg[c];
      idx += (u32)(npc < kr ? npc : kr);          // rounds 0 .. kr-1 of column c
      if (c < col && npc > kr) idx++;              // round kr of the columns in front
    }
    kl = (u32)(p.col_pairbeg[gf] - (i64)h.pair_base) + idx;
  }
  unsigned char* tb = p.blob + (size_t)h.blob16 * 16;
  const u32 po = off_pairs(h.cols_off, (u32)h.ncol);
  uint2 rec;
  rec.x = off8(find_slot(p, vin, col)) | (off8(find_slot(p, d[edge_of(P, I)] - 1, col)) << 8) |
          (off8(find_slot(p, d[edge_of(Q, I)] - 1, col)) << 16) | (off8(find_slot(p, d[edge_of(I, O)] - 1, col)) << 24);
  rec.y = (p.pair_io[k] & 0xffffffu) | (fl << 24);
  reinterpret_cast<uint2*>(tb + po)[kl] = rec;
  // mirror candidate (e_PQ, v_in); the first pair of a closed ring is completed by the last one and travels with the column
  const bool first_of_closed = p.col_closed[col] == 1 && k == p.col_pairbeg[col];
  const size_t m = (size_t)p.mir_base[tile] + 5u * (u32)h.ncol + kl;
  p.mkey[m] = first_of_closed ? NONE : gslot(p, col, vin);
  p.mval[m] = po / 8 + kl;
  if (p.end_slots) p.end_slots[k] = (fl & PF_END) ? gslot(p, col, d[O] - 1) : NONE;
}

__global__ void pack_cols(PackParams p) {
  i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (j >= p.ncols) return;
  if (p.col_closed[j] == 2) return;     // not an edge column
  const i64 kb = p.col_pairbeg[j], ke = p.col_pairbeg[j + 1];
  const u32 tile = p.col_tile[j];
  const TileHdr h = p.hdr[tile];
  const bool closed = p.col_closed[j] == 1;
  // reference orientation (P,Q) = first ring pair
  const u32 c0 = p.pair_code[kb];
  const i32* d0 = p.celldofs + (i64)p.pair_cell[kb] * 10;
  const i64 vP = d0[c0 & 3] - 1, vQ = d0[(c0 >> 2) & 3] - 1;
  // closing vertex: closed -> in-vertex of the first pair; open -> out-vertex of the last pair
  const i64 kc = closed ? kb : ke - 1;
  const u32 cc = p.pair_code[kc];
  const i32* dc = p.celldofs + (i64)p.pair_cell[kc] * 10;
  const int Pc = cc & 3, Qc = (cc >> 2) & 3, Vc = closed ? ((cc >> 4) & 3) : ((cc >> 6) & 3);
  const i64 vC = dc[Vc] - 1;
  uint4 a;
  a.x = (p.col_pq[j] & 0xffffffu) | ((u32)(ke - kb) << 24);
  a.y = off8(find_slot(p, vP, j)) | (off8(find_slot(p, vQ, j)) << 8) | (off8(find_slot(p, j, j)) << 16) | ((closed ? 1u : 0u) << 24);
  a.z = off8(find_slot(p, vC, j)) | (off8(find_slot(p, dc[edge_of(Pc, Vc)] - 1, j)) << 8) | (off8(find_slot(p, dc[edge_of(Qc, Vc)] - 1, j)) << 16);
  a.w = (p.col_abase[j] & 0xffffu) | ((u32)(kb - (i64)h.pair_base) << 16);
  const u32 cl = (u32)(j - h.c0);
  uint4* dst = reinterpret_cast<uint4*>(p.blob + (size_t)h.blob16 * 16 + h.cols_off) + 2 * cl;
  dst[0] = a; dst[1] = make_uint4(0, 0, 0, 0);
  // mirror candidates of the column: parked doubles A, B, q0, W at words 4 cl + {0,1,2,3} of the column section
  const size_t m = (size_t)p.mir_base[tile] + 5u * cl;
  const u32 w0 = h.cols_off / 8 + 4 * cl;
  p.mkey[m + 0] = gslot(p, j, vP); p.mval[m + 0] = w0 + 0;    // (e_PQ, v_P) <- A
  p.mkey[m + 1] = gslot(p, j, vQ); p.mval[m + 1] = w0 + 1;    // (e_PQ, v_Q) <- B
  p.mkey[m + 2] = gslot(p, j, vC); p.mval[m + 2] = w0 + 2;    // (e_PQ, v_closing) <- q0
  p.mkey[m + 3] = gslot(p, vQ, vP); p.mval[m + 3] = w0 + 3;   // (v_Q, v_P) <- W
  p.mkey[m + 4] = gslot(p, vP, vQ); p.mval[m + 4] = w0 + 3;   // (v_P, v_Q) <- W
}

// one block per tile: header, group table, node coordinates and the destination-sorted mirror list into the blob.  Only the valid
// mirror candidates are stored (they sort in front of NONE): destinations, then -- 16-byte aligned behind the LAST VALID one -- the
// sources; the tile directory gets the trimmed size, so the bulk load of the tile ends there (the blob keeps its place and its
// untrimmed allocation; ~11 % of the candidates of a uniform_refine grid are invalid = 1.4 % of the step's DRAM traffic).
__host__ __device__ inline u32 msrc_words(u32 nmir) { return (nmir + 3u) & ~3u; }     // u32 words between the two lists
__global__ void pack_tile_rest(const TileHdr* hdr, const uint4* groups, int nw, const u32* tile_nodeids, const double* coords,
                               const int* mir_base, const u32* mkey_sorted, const u32* mval_sorted, unsigned char* blob, uint2* tile_dir) {
  __shared__ int s_valid;
  const TileHdr h = hdr[blockIdx.x];
  unsigned char* tb = blob + (size_t)h.blob16 * 16;
  if (threadIdx.x == 0) s_valid = 0;
  __syncthreads();
  if ((int)threadIdx.x <= nw) reinterpret_cast<uint4*>(tb + 48)[threadIdx.x] = groups[(size_t)blockIdx.x * (nw + 1) + threadIdx.x];
  double* X = reinterpret_cast<double*>(tb + off_xyz(h.cols_off, (u32)h.ncol, (u32)h.npairs));
  for (int i = threadIdx.x; i < h.nnodes; i += blockDim.x) {
    const double* xg = coords + (size_t)(tile_nodeids[(size_t)h.node_base + i] - 1) * 3;
    X[3 * i] = xg[0]; X[3 * i + 1] = xg[1]; X[3 * i + 2] = xg[2];
  }
  const int m0 = mir_base[blockIdx.x], nm = mir_base[blockIdx.x + 1] - m0;
  int mine = 0;
  for (int i = threadIdx.x; i < nm; i += blockDim.x) mine += mkey_sorted[m0 + i] != NONE;
  atomicAdd(&s_valid, mine);
  __syncthreads();
  const int nmir = s_valid;
  const u32 md_off = off_mdst(h.cols_off, (u32)h.ncol, (u32)h.npairs, (u32)h.nnodes);
  u32* md = reinterpret_cast<u32*>(tb + md_off);
  unsigned short* ms = reinterpret_cast<unsigned short*>(md + msrc_words((u32)nmir));
  for (int i = threadIdx.x; i < nmir; i += blockDim.x) {
    md[i] = mkey_sorted[m0 + i]; ms[i] = (unsigned short)mval_sorted[m0 + i];
  }
  if (threadIdx.x == 0) {
    TileHdr hb = h;
    hb.node_base = (u32)nmir;         // nmir: the valid candidates sort in front of NONE
    hb.blob_bytes = md_off + 4u * msrc_words((u32)nmir) + pad16(2u * (u32)nmir);
    *reinterpret_cast<TileHdr*>(tb) = hb;
    tile_dir[blockIdx.x].y = hb.blob_bytes;
  }
}

// the grid's coordinates changed (grmp_grid_update_geometry): refresh the tile-blocked copies inside the blobs
__global__ void repack_tile_coords(const TileHdr* hdr, const u32* tile_nodeids, const double* coords, unsigned char* blob) {
  const TileHdr h = hdr[blockIdx.x];
  double* X = reinterpret_cast<double*>(blob + (size_t)h.blob16 * 16 + off_xyz(h.cols_off, (u32)h.ncol, (u32)h.npairs));
  for (int i = threadIdx.x; i < h.nnodes; i += blockDim.x) {
    const double* xg = coords + (size_t)(tile_nodeids[(size_t)h.node_base + i] - 1) * 3;
    X[3 * i] = xg[0]; X[3 * i + 1] = xg[1]; X[3 * i + 2] = xg[2];
  }
}

struct EdgeParams {
  const uint2* tile_dir;    // per tile: blob offset (16-byte units), blob bytes
  const unsigned char* blob;
  const u32* end_slots;     // [npairs] or null (only partitions have multi-chain columns)
  double factor;
  double* nzval;
  i64 halo_first;           // first slot of the halo columns: group ranges are stored up to here only
  unsigned long long* tile_counter;   // dynamic tile scheduler: running claim counter.  Every CTA claims until its first miss, so one launch
                            // advances it by exactly claims_per_step = ntiles + gridDim.x and claim c belongs to tile c % claims_per_step:
                            // no reset between launches, which lets the next launch start (programmatic dependent launch) before
                            // the diagonal kernel of this one has finished
  u32 claims_per_step;
  int ntiles;
  u32 in_stride;            // bytes of one input buffer (largest blob)
  u32 slot_elems;           // doubles per warp stage slot
  int nbuf;                 // depth of the input ring (2 or 3)
  unsigned long long* prof; // GRMP_FAST_PROF: 8 cycle counters per CTA (service warp: wait done, mirror write-out, total; consumer warp 0: wait full, total)
  int dbg;                  // GRMP_DEBUG_FLAGS (timing experiments only): 1 skip the mirror write-out, 2 skip the diagonal kernel, 4 skip the ring
                            // walk, 8 skip the bulk stores, 16 mirror stores without the evict-last hint, 32 / 64 evict-first hint on the
                            // bulk stores / blob loads
};

// programmatic dependent launch (no-ops when the kernel was launched without the attribute)
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ double fast_rcp(double d) {   // 1/d to ~1 ulp for normal d: MUFU.RCP64H + two Newton steps
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double t = fma(-d, r, 1.0);
  r = fma(r, t, r);
  t = fma(-d, r, 1.0);
  r = fma(r, t, r);
  return r;
}
__device__ __forceinline__ u64 l2_policy_evict_first() {
  u64 pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ u64 l2_policy_evict_last() {
  u64 pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_hint(double* ptr, double v, u64 pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(ptr), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void tile_load(const EdgeParams& p, uint2 dir, unsigned dst_smem, unsigned mbar_a, u64 pol) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(dir.y) : "memory");
  if (!(p.dbg & 64))
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(p.blob + (size_t)dir.x * 16), "r"(dir.y), "r"(mbar_a)
                 : "memory");
  else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
                 "l"(p.blob + (size_t)dir.x * 16), "r"(dir.y), "r"(mbar_a), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar_a, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(mbar_a), "r"(parity)
      : "memory");
}

// service warp: the tile's parked mirror values -> their slots in the vertex columns, in destination order
__device__ __forceinline__ void mirror_writeout(const EdgeParams& p, const unsigned char* in, int lane, int nlanes) {
  const int4 h0 = reinterpret_cast<const int4*>(in)[0], h2 = reinterpret_cast<const int4*>(in)[2];
  const u32 ncol = (u32)h0.y, nnodes = (u32)h0.z, npairs = (u32)h0.w, nmir = (u32)h2.y, cols_off = (u32)h2.w;
  const u32* __restrict__ md = reinterpret_cast<const u32*>(in + off_mdst(cols_off, ncol, npairs, nnodes));
  const unsigned short* __restrict__ ms = reinterpret_cast<const unsigned short*>(md + msrc_words(nmir));
  const double* __restrict__ words = reinterpret_cast<const double*>(in);
  const u64 pol_keep = l2_policy_evict_last();
  u32 i = lane;
#ifndef GRMP_WRITEOUT_U
#define GRMP_WRITEOUT_U 8
#endif
  constexpr int U = GRMP_WRITEOUT_U;                   // independent gather/store chains per lane
  for (; i + (U - 1) * nlanes < nmir; i += U * nlanes) {
    u32 d[U], w[U];
    double v[U];
#pragma unroll
    for (int u = 0; u < U; u++) { d[u] = md[i + u * nlanes]; w[u] = ms[i + u * nlanes]; }
#pragma unroll
    for (int u = 0; u < U; u++) v[u] = words[w[u]];
    // evict-last: the sectors of a vertex column are completed by other tiles later; keeping the partially written lines in
    // L2 until then (the streaming loads / bulk stores are evict-first) is worth 8 % of the kernel
    if (p.dbg & 16) {
#pragma unroll
      for (int u = 0; u < U; u++) p.nzval[d[u]] = v[u];
    } else {
#pragma unroll
      for (int u = 0; u < U; u++) st_hint(p.nzval + d[u], v[u], pol_keep);
    }
  }
  for (; i < nmir; i += nlanes) st_hint(p.nzval + md[i], words[ms[i]], pol_keep);
}

// Persistent CTAs of NW consumer warps + 1 service warp.  Tiles are claimed from a global counter.  The service warp keeps the
// 2-deep input ring filled (TMA bulk loads, full[] mbarriers) and, once all consumer warps have left a tile (done[] mbarriers),
// writes its mirror values out and reuses the buffer.  Consumer warp w owns group w of every tile, stages the group's nzval
// range in its own slot and stores it with its own TMA bulk store.  No CTA-wide barrier after the set-up.
template <int NW>
__global__ void __launch_bounds__((NW + NSVC) * 32, (NW >= 5 ? 2 : NW == 4 ? 3 : 4)) p2tet_edge_kernel(const EdgeParams p) {
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ __align__(8) unsigned long long mbar[6];      // full[0..2], done[0..2]
  __shared__ int s_tile[3];                                // tile in each input buffer, -1 = no more tiles
  __shared__ int s_next;                                   // service warps: tile claimed for the next load
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const unsigned full_a = (unsigned)__cvta_generic_to_shared(&mbar[0]), done_a = full_a + 24;
  const int NB = p.nbuf;
  const unsigned in_a = (unsigned)__cvta_generic_to_shared(smraw);
  if (tid == 0) {
    for (int b = 0; b < 3; b++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full_a + 8 * b));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(done_a + 8 * b), "r"(NW));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // the diagonal kernel may become resident as the CTAs of this grid retire; it waits for the whole grid before it reads
    pdl_launch_dependents();
  }
  __syncthreads();
  // Until pdl_wait() this kernel only reads what no earlier kernel of the stream writes after the symbolic pass (blobs, tile
  // directory) and claims tiles: with a programmatic launch the first tile is loaded and walked while the previous step's
  // diagonal kernel drains.  Every thread passes pdl_wait() before its first global store.
  bool waited = false;
  if (warp >= NW) {   // ---- service warps ----
    const u64 pol_stream = l2_policy_evict_first();
    const int slane = tid - NW * 32;                        // 0 .. 32 NSVC - 1; thread 0 of the service warps loads tiles
    int it = 0;
    int b = 0, use = 0;                                       // buffer of iteration it, how often it has been filled before
    long long c_wait = 0, c_wo = 0;
    const long long c_start = clock64();
    for (;; it++) {
      int t = 0;
      uint2 dir = make_uint2(0, 0);
      if (slane == 0) {
        // tiles are claimed dynamically: SMs do not run at the same speed, a static split leaves the slowest SM as the tail
        t = (int)(atomicAdd(p.tile_counter, 1ull) % p.claims_per_step);
        if (t < p.ntiles) dir = __ldg(p.tile_dir + t);
      }
      if (use >= 1) {
        const long long c0 = clock64();
        mbar_wait(done_a + 8 * b, (unsigned)(use - 1) & 1u);   // all consumer warps have left the tile in this buffer
        const long long c1 = clock64();
        if (!waited) { pdl_wait(); waited = true; }
        if (!(p.dbg & 1)) mirror_writeout(p, smraw + (size_t)b * p.in_stride, slane, 32 * NSVC);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // parked values (generic writes) before the next bulk load
        c_wait += c1 - c0; c_wo += clock64() - c1;
      }
      if (slane == 0) s_next = t;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * NSVC) : "memory");     // write-out finished by all service warps; s_next visible
      t = *reinterpret_cast<volatile int*>(&s_next);
      asm volatile("bar.sync 1, %0;" ::"n"(32 * NSVC) : "memory");     // everybody has read s_next
      if (t >= p.ntiles) {
        if (slane == 0) {
          s_tile[b] = -1;
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_a + 8 * b) : "memory");
        }
        break;
      }
      if (slane == 0) {
        s_tile[b] = t;
        tile_load(p, dir, in_a + b * p.in_stride, full_a + 8 * b, pol_stream);
      }
      if (++b == NB) { b = 0; use++; }
    }
    // the other buffers still hold the last NB - 1 tiles (iterations it - NB + 1 .. it - 1)
    for (int j = (it >= NB - 1 ? it - (NB - 1) : 0); j < it; j++) {
      const int b2 = j % NB;
      mbar_wait(done_a + 8 * b2, (unsigned)(j / NB) & 1u);
      if (!waited) { pdl_wait(); waited = true; }
      if (!(p.dbg & 1)) mirror_writeout(p, smraw + (size_t)b2 * p.in_stride, slane, 32 * NSVC);
    }
    if (!waited) pdl_wait();
    if (p.prof != nullptr && slane == 0) {
      unsigned long long* q = p.prof + 8 * (size_t)blockIdx.x;
      q[0] = (unsigned long long)c_wait; q[1] = (unsigned long long)c_wo; q[2] = (unsigned long long)(clock64() - c_start); q[7] = (unsigned long long)it;
    }
    return;
  }
  // ---- consumers ----
  const u64 pol_stream = l2_policy_evict_first();
  double* const slot = reinterpret_cast<double*>(smraw + (size_t)NB * p.in_stride) + (size_t)warp * p.slot_elems;
  long long cc_wait = 0;
  const long long cc_start = clock64();
  for (int it = 0, cur = 0, use = 0;; it++) {
    unsigned char* in = smraw + (size_t)cur * p.in_stride;
    {
      const long long c0 = clock64();
      mbar_wait(full_a + 8 * cur, (unsigned)use & 1u);
      cc_wait += clock64() - c0;
    }
    if (*reinterpret_cast<volatile int*>(&s_tile[cur]) < 0) break;
    const int4 h0 = reinterpret_cast<const int4*>(in)[0], h1 = reinterpret_cast<const int4*>(in)[1], h2 = reinterpret_cast<const int4*>(in)[2];
    const uint4 gr0 = reinterpret_cast<const uint4*>(in + 48)[warp], gr1 = reinterpret_cast<const uint4*>(in + 48)[warp + 1];
    const int col = (int)gr0.x + lane;                      // tile-local column of this lane
    const bool has_col = col < (int)gr1.x && !(p.dbg & 4);
    const i64 g0 = ((i64)(u32)h1.x | ((i64)h1.y << 32)) + gr0.y;   // first nzval slot of the group
    // halo columns (a suffix of the slot range) are walked for their mirrors only: nothing of them is stored
    const int nnz_w = (int)max((i64)0, min((i64)(gr1.y - gr0.y), p.halo_first - g0));
    const u32 cols_off = (u32)h2.w, pairs_off = off_pairs(cols_off, (u32)h0.y), xyz_off = off_xyz(cols_off, (u32)h0.y, (u32)h0.w);
    const double* __restrict__ X = reinterpret_cast<const double*>(in + xyz_off);
    // stage[i] mirrors nzval[g0 + i]; it is shifted by one element when g0 is odd so that shared and global addresses of the
    // same element are 16-byte aligned together (TMA bulk store).  Every slot is written exactly once -> no zero-init.
    const int odd = (int)(g0 & 1);
    double* __restrict__ stage = slot + odd;
    // this warp's previous bulk store must have finished reading the slot before it is overwritten
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    {
      uint4 ca = make_uint4(0, 0, 0, 0);
      if (has_col) ca = *reinterpret_cast<const uint4*>(in + cols_off + 32u * (u32)col);
      const u32 np = ca.x >> 24;                            // 0 for lanes without a column
      const u32 maxnp = __reduce_max_sync(0xffffffffu, np);
      double mA = 0.0, mB = 0.0, mW = 0.0, mq0 = 0.0;
      // pair records of the group, round-major and compact: record of (lane, round j) = gp[base_j + rank of the lane among the
      // lanes with np > j].  next_rec(j) must be called for j = 0, 1, 2, ... by the whole warp.
      const uint2* gp = reinterpret_cast<const uint2*>(in + pairs_off) + gr0.z;
      double* gpark = reinterpret_cast<double*>(in + pairs_off) + gr0.z;    // consumed pair records take the mirror values
      const u32 lt = (1u << lane) - 1u;
      u32 rbase = 0;
      auto next_rec = [&](u32 j) -> u32 {
        const u32 bal = __ballot_sync(0xffffffffu, j < np);
        const u32 idx = rbase + __popc(bal & lt);
        rbase += __popc(bal);
        return idx;
      };
      u32 i0 = next_rec(0), i1 = next_rec(1);
      double* __restrict__ a = stage + (ca.w & 0xffffu);
      const bool closed = (ca.y >> 24) & 1u;
      const u32 lp = ca.x & 0xfffu, lq = (ca.x >> 12) & 0xfffu;
      const double px = X[3 * lp], py = X[3 * lp + 1], pz = X[3 * lp + 2];
      const double ax = X[3 * lq] - px, ay = X[3 * lq + 1] - py, az = X[3 * lq + 2] - pz;
      const double c8 = p.factor * (0.8 / 6.0);      // S' = 0.8 * S = c8 / |det| * n_a.n_b
      uint2 r0 = make_uint2(0, 0), r1 = make_uint2(0, 0);
      if (np > 0) r0 = gp[i0];
      if (np > 1) r1 = gp[i1];
      double bx, by, bz, mcx, mcy, mcz;               // b = c_in - p, mc = a x b (carried around the ring)
      {
        const u32 li = r0.y & 0xfffu;
        bx = X[3 * li] - px; by = X[3 * li + 1] - py; bz = X[3 * li + 2] - pz;
        mcx = ay * bz - az * by; mcy = az * bx - ax * bz; mcz = ax * by - ay * bx;
      }
      double ox, oy, oz;                              // coordinates of the out-vertex of the current pair
      {
        const u32 lo = (r0.y >> 12) & 0xfffu;
        ox = X[3 * lo]; oy = X[3 * lo + 1]; oz = X[3 * lo + 2];
      }
      double R1 = 0.0, R2 = 0.0, R3 = 0.0;            // ring sums of S'_pq, S'_pi + S'_po, S'_qi + S'_qo
      double c0r = 0.0, c1r = 0.0, c2r = 0.0;          // carry: partial rows of the shared ring vertex
      double f0 = 0.0, f1 = 0.0, f2 = 0.0;             // closed ring: in-rows of the first pair, completed at the end
#pragma unroll 2
      for (u32 k = 0; k < maxnp; k++) {
        const u32 i2 = next_rec(k + 2);
        if (k < np) {
          // software pipeline: out-vertex of the next pair, record after next
          const u32 ln = (r1.y >> 12) & 0xfffu;
          const double nx = X[3 * ln], ny = X[3 * ln + 1], nz = X[3 * ln + 2];
          uint2 r2 = r1;
          if (k + 2 < np) r2 = gp[i2];
          const u32 fl = r0.y >> 24;
          if (fl & PF_RESET) {
            const u32 li = r0.y & 0xfffu;
            bx = X[3 * li] - px; by = X[3 * li + 1] - py; bz = X[3 * li + 2] - pz;
            mcx = ay * bz - az * by; mcy = az * bx - ax * bz; mcz = ax * by - ay * bx;
            c0r = 0.0; c1r = 0.0; c2r = 0.0;
          }
          // cell (p, q, in, out): a = q-p, b = in-p, e = out-p;  n_q = b x e, n_in = -(a x e), n_out = a x b, n_p = -(n_q+n_in+n_out)
          const double ex = ox - px, ey = oy - py, ez = oz - pz;
          const double mdx = ay * ez - az * ey, mdy = az * ex - ax * ez, mdz = ax * ey - ay * ex;
          const double gx = by * ez - bz * ey, gy = bz * ex - bx * ez, gz = bx * ey - by * ex;
          const double det = ax * gx + ay * gy + az * gz;
          const double s = c8 * fast_rcp(fabs(det));
          const double npx = mdx - gx - mcx, npy = mdy - gy - mcy, npz = mdz - gz - mcz;
          const double spq = s * (npx * gx + npy * gy + npz * gz);
          const double spi = -s * (npx * mdx + npy * mdy + npz * mdz);
          const double spo = s * (npx * mcx + npy * mcy + npz * mcz);
          const double sqi = -s * (gx * mdx + gy * mdy + gz * mdz);
          const double sqo = s * (gx * mcx + gy * mcy + gz * mcz);
          // with S' = 0.8 S and the zero row sums of S (S_pp = -(S_pq+S_pi+S_po), S_qq likewise):
          //   v_in: -0.2(S_pi+S_qi)   e_P,in: 0.8(2 S_qi - S_po)   e_Q,in: 0.8(2 S_pi - S_qo)   e_in,out: 0.8(S_pi+S_po+S_qi+S_qo)
          const double tp = spi + spo, tq = sqi + sqo;
          R1 += spq; R2 += tp; R3 += tq;
          const double in0 = c0r - 0.25 * (spi + sqi);
          const double in1 = c1r + (2.0 * sqi - spo);
          const double in2 = c2r + (2.0 * spi - sqo);
          c0r = -0.25 * (spo + sqo);
          c1r = 2.0 * sqo - spi;
          c2r = 2.0 * spo - sqi;
          const u32 o3 = r0.x >> 24;
          if (o3 != 255u) a[o3] = tp + tq;
          if ((fl & PF_END) && p.end_slots != nullptr && !(p.dbg & 1)) {     // chain end inside a halo column: mirror (e_PQ, v_out)
            const u32 es = p.end_slots[(size_t)(u32)h2.x + (ca.w >> 16) + k];
            if (es != NONE) p.nzval[es] = c0r;
          }
          if (closed && k == 0) {
            f0 = in0; f1 = in1; f2 = in2;                                  // partner is the last pair of the ring
          } else {
            const u32 o0 = r0.x & 255u, o1 = (r0.x >> 8) & 255u, o2 = (r0.x >> 16) & 255u;
            if (o0 != 255u) a[o0] = in0;
            if (o1 != 255u) a[o1] = in1;
            if (o2 != 255u) a[o2] = in2;
            gpark[i0] = in0;                                               // mirror (e_PQ, v_in)
          }
          bx = ex; by = ey; bz = ez; mcx = mdx; mcy = mdy; mcz = mdz;
          ox = nx; oy = ny; oz = nz;
          r0 = r1; r1 = r2;
        }
        i0 = i1; i1 = i2;
      }
      if (np > 0) {
        // closing rows: closed ring -> first pair's in-rows + last carry; open chain -> last pair's out-rows
        {
          const double q0 = closed ? f0 + c0r : c0r, q1 = closed ? f1 + c1r : c1r, q2 = closed ? f2 + c2r : c2r;
          const u32 o0 = ca.z & 255u, o1 = (ca.z >> 8) & 255u, o2 = (ca.z >> 16) & 255u;
          if (o0 != 255u) a[o0] = q0;
          if (o1 != 255u) a[o1] = q1;
          if (o2 != 255u) a[o2] = q2;
          mq0 = q0;
        }
        {
          // rows v_P, v_Q, e_PQ and the (v_P, v_Q) coupling from the ring sums (S'_pp = -(S'_pq + S'_pi + S'_po)):
          //   A = sum 0.6 S_pq - 0.2 S_pp = R1 + R2/4,  C = 1.6 sum (S_pp+S_qq+S_pq) = -2 (R1+R2+R3),  W = -0.2 sum S_pq = -R1/4
          const double A = R1 + 0.25 * R2, B = R1 + 0.25 * R3, C = -2.0 * (R1 + R2 + R3), W = -0.25 * R1;
          const u32 oA = ca.y & 255u, oB = (ca.y >> 8) & 255u, oC = (ca.y >> 16) & 255u;
          if (oA != 255u) a[oA] = A;
          if (oB != 255u) a[oB] = B;
          if (oC != 255u) a[oC] = C;
          mA = A; mB = B; mW = W;
        }
      }
      if (has_col) {
        // park the column's mirror values in its own (consumed) record
        double2* rec = reinterpret_cast<double2*>(in + cols_off + 32u * (u32)col);
        rec[0] = make_double2(mA, mB); rec[1] = make_double2(mq0, mW);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // stage writes -> visible to the bulk store
    __syncwarp();
    if (!waited) { pdl_wait(); waited = true; }
    {
      // the group's nzval range is contiguous: one TMA bulk store (cp.async.bulk shared -> global) of the 16-byte aligned
      // body, the (at most one) unaligned element at either end by ordinary stores
      double* __restrict__ dst = p.nzval + g0;
      const int i0 = odd;                                   // first element whose address is 16-byte aligned
      const int nb = (nnz_w > i0) ? ((nnz_w - i0) & ~1) : 0;  // elements in the bulk body
      if (lane == 0) {
        if (nb > 0 && !(p.dbg & 8)) {
          const unsigned src = (unsigned)__cvta_generic_to_shared(stage + i0);
          if (!(p.dbg & 32)) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + i0), "r"(src), "r"(nb * 8) : "memory");
          else asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst + i0), "r"(src), "r"(nb * 8), "l"(pol_stream) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        // every lane is done with this input buffer (ordered by the __syncwarp above); release makes the parked values visible
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(done_a + 8 * cur) : "memory");
      }
      if (lane == 1 && i0 == 1 && nnz_w > 0) dst[0] = stage[0];
      if (lane == 2 && i0 + nb < nnz_w) dst[i0 + nb] = stage[i0 + nb];
    }
    if (++cur == NB) { cur = 0; use++; }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem must stay alive until read
  if (!waited) pdl_wait();
  if (p.prof != nullptr && tid == 0) {
    unsigned long long* q = p.prof + 8 * (size_t)blockIdx.x;
    q[3] = (unsigned long long)cc_wait; q[4] = (unsigned long long)(clock64() - cc_start);
  }
}

// Diagonal of the vertex columns.  The vertex-vertex block of the P2 stiffness matrix is the P1 stiffness matrix K1 scaled
// entrywise (A[v,v] = 0.6 K1[v,v], A[w,v] = -0.2 K1[w,v]: the local vertex-vertex block is 0.6 / -0.2 S), and the columns of
// K1 sum to zero (the P1 basis is a partition of unity), hence A[v,v] = 3 * sum_{w != v} A[w,v] over the VERTEX rows w of the
// column only -- ~15 of its ~65 entries, every one of them mirrored into place by the edge kernel.  Columns with more than 128
// entries (vertices of very high valence) use the plain column sum instead: A[v,v] = -sum_{i != v} A[i,v] (P2 partition of
// unity).  Eight lanes per column, fixed shuffle tree -> deterministic.
// vrec: {diagonal slot | NONE, first slot, #slots, mode}, {row-is-vertex mask x4}.
__global__ void __launch_bounds__(256) p2tet_vertex_diag_kernel(const uint4* __restrict__ vrec, i64 nv, double* nzval) {
  const i64 w = (blockIdx.x * (i64)blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  // once the last wave of this grid is resident the next step's edge kernel may start loading its first tiles
  if (threadIdx.x == 0) pdl_launch_dependents();
  uint4 r = make_uint4(NONE, 0, 0, 0), m = make_uint4(0, 0, 0, 0);
  if (w < nv) { r = __ldg(vrec + 2 * w); m = __ldg(vrec + 2 * w + 1); }      // records of the symbolic pass: not written by the edge kernel
  pdl_wait();                                                                 // the edge kernel's stores (all of nzval) are complete and visible
  double s = 0.0;
  if (r.x != NONE) {
    const double* __restrict__ c = nzval + r.y;
    const u32 d = r.x - r.y;
    if (r.w == 0) {
      // the vertex rows are the lowest row indices of the column on an unpartitioned grid: usually only mask word 0 is set
#pragma unroll
      for (int wd = 0; wd < 4; wd++) {
        const u32 word = wd == 0 ? m.x : (wd == 1 ? m.y : (wd == 2 ? m.z : m.w));
        if (word != 0u) {
          double v[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {       // slots 32 wd + sub + 8 u: independent loads
            const u32 k = 32 * wd + sub + 8 * u;
            const bool take = k != d && ((word >> (sub + 8 * u)) & 1u);   // mask bits exist only for slots < #slots
            v[u] = take ? c[k] : 0.0;
          }
          s += (v[0] + v[1]) + (v[2] + v[3]);
        }
      }
      s *= 3.0;
    } else {
      for (u32 k = sub; k < r.z; k += 8) s -= (k == d) ? 0.0 : c[k];
    }
  }
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (sub == 0 && r.x != NONE) nzval[r.x] = s;
}

// col_kind[row] == 2 marks vertex dofs
__global__ void find_diag_slots(const u32* vcols, i64 nv, const i64* colptr, const i64* rowval, const unsigned char* col_kind, uint4* vrec) {
  i64 w = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (w >= nv) return;
  const i64 col = vcols[w];
  const i64 beg = colptr[col] - 1, end = colptr[col + 1] - 1;
  const bool masked = end - beg <= 128;
  u32 d = NONE, mask[4] = {0, 0, 0, 0};
  for (i64 k = beg; k < end; k++) {
    const i64 row = rowval[k] - 1;
    if (row == col) d = (u32)k;
    else if (masked && col_kind[row] == 2) mask[(k - beg) >> 5] |= 1u << ((k - beg) & 31);
  }
  vrec[2 * w] = make_uint4(d, (u32)beg, (u32)(end - beg), masked ? 0u : 1u);
  vrec[2 * w + 1] = make_uint4(mask[0], mask[1], mask[2], mask[3]);
}

// ==== device-side build (J3): ring order of every edge column and tile packing on the GPU ==========================================
// The host build below (GRMP_FAST_HOST_BUILD=1) is kept for cross-validation: both produce the same ring orders and the same matrix
// bit for bit; the device build cuts tiles additionally at chunk boundaries (chunks of TB_CHUNK columns are packed independently).
constexpr int MAXRING = 64;          // cells around one edge that the device ring ordering handles (host build: 255)
constexpr int TB_CHUNK = 16384;      // columns per independently packed chunk (~70 tiles)
constexpr int TB_HASH = 4096;        // open-addressing node set of the tile being filled; the device build is used when a blob holds < TB_HASH / 2 nodes
enum { CK_OPEN = 0, CK_CLOSED = 1, CK_VERTEX = 2 };

struct RingParams {
  const u32* gcell; const u32* gsrc; const i64* pairbeg; const i64* colptr; const i32* cellnodes;
  i64 ncells, ncols, ncols_owned;
  u32* pair_cell; u32* pair_code; u32* col_of_pair; i32* pair_in; i32* pair_out; i32* col_P; i32* col_Q; unsigned char* col_closed;
  int* flags;      // [0] error code, [1] a multi-chain column exists (PF_END)
};

// thread per column: the host loop (2a) below, line for line
__global__ void __launch_bounds__(128) ring_order_kernel(const RingParams p) {
  const i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (j >= p.ncols) return;
  const i64 kb = p.pairbeg[j], ke = p.pairbeg[j + 1];
  p.col_closed[j] = CK_VERTEX;
  if (ke == kb) return;
  const i64 len = p.colptr[j + 1] - p.colptr[j];
  const int lj0 = (int)(p.gsrc[kb] / (u32)p.ncells);
  if (lj0 < 4) {   // vertex column: filled by mirrors + the diagonal kernel
    for (i64 k = kb; k < ke; k++) { p.pair_cell[k] = p.gcell[k]; p.pair_code[k] = 0; p.col_of_pair[k] = (u32)j; }
    return;
  }
  if (len > 254 || ke - kb > MAXRING) { atomicMax(p.flags, ke - kb > MAXRING ? 7 : 1); return; }
  const int n = (int)(ke - kb);
  u32 rc[MAXRING]; unsigned char rl[MAXRING]; i32 nR[MAXRING], nS[MAXRING];    // cell, local ids P | Q<<2 | R<<4 | S<<6, ring nodes
  i32 P0 = 0, Q0 = 0;
  for (int t = 0; t < n; t++) {
    const u32 c = p.gcell[kb + t];
    const int lj = (int)(p.gsrc[kb + t] / (u32)p.ncells);
    if (lj < 4) { atomicMax(p.flags, 2); return; }
    int pl, ql; edge_nodes(lj - 4, pl, ql);
    int rloc = -1, sloc = -1;
    for (int v = 0; v < 4; v++) if (v != pl && v != ql) { if (rloc < 0) rloc = v; else sloc = v; }
    const int4 cn4 = __ldg(reinterpret_cast<const int4*>(p.cellnodes) + c);
    const i32 cn[4] = {cn4.x, cn4.y, cn4.z, cn4.w};
    if (t == 0) { P0 = cn[pl]; Q0 = cn[ql]; }
    if (cn[pl] != P0) { const int tmp = pl; pl = ql; ql = tmp; }
    if (cn[pl] != P0 || cn[ql] != Q0) { atomicMax(p.flags, 3); return; }
    rc[t] = c; rl[t] = (unsigned char)(pl | (ql << 2) | (rloc << 4) | (sloc << 6)); nR[t] = cn[rloc]; nS[t] = cn[sloc];
  }
  p.col_P[j] = P0; p.col_Q[j] = Q0;
  bool closed = true;
  u64 deg1R = 0, deg1S = 0;       // bit t: the R / S node of cell t has degree 1
  for (int t = 0; t < n; t++) {
    int dR = 0, dS = 0;
    for (int u = 0; u < n; u++) {
      dR += (nR[u] == nR[t]) + (nS[u] == nR[t]);
      dS += (nR[u] == nS[t]) + (nS[u] == nS[t]);
    }
    if (dR > 2 || dS > 2) { atomicMax(p.flags, 4); return; }
    if (dR == 1) deg1R |= 1ull << t;
    if (dS == 1) deg1S |= 1ull << t;
    if (dR == 1 || dS == 1) closed = false;
  }
  u64 used = 0;
  int step = 0, nchains = 0;
  while (step < n) {
    int start = -1, start_in_is_R = 1;
    if (closed) { if (step != 0) { atomicMax(p.flags, 5); return; } start = 0; }
    else
      for (int t = 0; t < n && start < 0; t++)
        if (!((used >> t) & 1ull)) { if ((deg1R >> t) & 1ull) { start = t; start_in_is_R = 1; } else if ((deg1S >> t) & 1ull) { start = t; start_in_is_R = 0; } }
    if (start < 0) { atomicMax(p.flags, 6); return; }
    if (nchains > 0 && j < p.ncols_owned) { atomicMax(p.flags, 8); return; }
    int curp = start;
    const i32 first_in = start_in_is_R ? nR[start] : nS[start];
    i32 vin = first_in;
    bool chain_first = true;
    while (true) {
      used |= 1ull << curp;
      const bool inR = (nR[curp] == vin);
      const int Pl = rl[curp] & 3, Ql = (rl[curp] >> 2) & 3, Rl = (rl[curp] >> 4) & 3, Sl = (rl[curp] >> 6) & 3;
      const int I = inR ? Rl : Sl, O = inR ? Sl : Rl;
      const i32 vout = inR ? nS[curp] : nR[curp];
      const i64 k = kb + step;
      u32 fl = 0u;
      if (chain_first && nchains > 0) fl |= PF_RESET;
      p.pair_cell[k] = rc[curp];
      u32 code = (u32)(Pl | (Ql << 2) | (I << 4) | (O << 6)) | (fl << 8);
      p.col_of_pair[k] = (u32)j;
      p.pair_in[k] = vin; p.pair_out[k] = vout;
      step++; chain_first = false;
      int nxt = -1;
      for (int u = 0; u < n; u++) if (!((used >> u) & 1ull) && (nR[u] == vout || nS[u] == vout)) { nxt = u; break; }
      if (nxt < 0) {
        if (closed && (step != n || vout != first_in)) { atomicMax(p.flags, 9); return; }
        if (!closed && step < n) { code |= (PF_END << 8); p.flags[1] = 1; }
        p.pair_code[k] = code;
        break;
      }
      p.pair_code[k] = code;
      vin = vout; curp = nxt;
    }
    nchains++;
  }
  p.col_closed[j] = closed ? CK_CLOSED : CK_OPEN;
}

struct TileBuildParams {
  const i64* pairbeg; const i64* colptr; const unsigned char* col_closed; const i32* pair_in; const i32* pair_out; const i32* col_P; const i32* col_Q;
  i64 ncols;
  int nw; i64 slot_cap, blob_cap; u32 cols_off;
  // phase 1 outputs / phase 2 inputs, per chunk
  int* chunk_ntiles; int* chunk_nnodes;       // counts (phase 1), exclusive offsets (phase 2)
  int emit;
  // phase 2 outputs
  TileHdr* hdr; uint4* groups; u32* tile_nodeids; u32* col_tile; u32* col_pq; u32* col_abase; u32* col_group; u32* col_gcount; u32* pair_io;
  int* flags;       // [2] max group nnz, [3] error
};

// one thread per chunk of TB_CHUNK columns: the host loop (2b) below with the tile's node set in a shared-memory hash table.
// Phase 1 (emit = 0) counts tiles and tile nodes of the chunk, phase 2 (emit = 1) writes everything at the scanned offsets.
__global__ void __launch_bounds__(32) tile_build_kernel(const TileBuildParams p) {
  __shared__ u32 h_node[TB_HASH];      // node id
  __shared__ u32 h_tag[TB_HASH];       // tile generation << 12 | tile-local id
  if (threadIdx.x != 0) return;
  const i64 chunk = blockIdx.x;
  const i64 j0 = chunk * TB_CHUNK, j1 = min(j0 + (i64)TB_CHUNK, p.ncols);
  for (int i = 0; i < TB_HASH; i++) { h_node[i] = 0xffffffffu; h_tag[i] = 0xffffffffu; }
  const int NW = p.nw;
  int tile_base = 0, node_base = 0;
  if (p.emit) { tile_base = p.chunk_ntiles[chunk]; node_base = p.chunk_nnodes[chunk]; }
  int cur_tile = 0;                        // chunk-local tile index = hash generation
  int cur_cols = 0, cur_nodes = 0, cur_groups = 0, grp_cols = 0;
  i64 grp_nnz = 0, cur_nnz = 0, cur_pairs = 0, tile_first_col = 0, tile_pair_base = 0, grp_first_col = 0;
  int nodes_total = 0, tile_node_base = 0, max_slot = 0;
  uint4 cur_grp[9];
  auto lookup = [&](u32 v, bool insert) -> int {      // tile-local id of node v in the open tile, -1 if absent (inserted when asked)
    u32 h = (v * 2654435761u) & (TB_HASH - 1);
    for (;;) {
      const bool live = h_node[h] != 0xffffffffu && (h_tag[h] >> 12) == (u32)cur_tile;
      if (live && h_node[h] == v) return (int)(h_tag[h] & 0xfffu);
      if (!live) {
        if (!insert) return -1;
        h_node[h] = v; h_tag[h] = ((u32)cur_tile << 12) | (u32)cur_nodes;
        if (p.emit) p.tile_nodeids[(size_t)node_base + nodes_total] = v;
        nodes_total++;
        return cur_nodes++;
      }
      h = (h + 1) & (TB_HASH - 1);
    }
  };
  auto close_group = [&](i64 end_col) {
    if (!p.emit) return;
    for (i64 c = grp_first_col; c < end_col; c++) { p.col_group[c] = (u32)grp_first_col; p.col_gcount[c] = (u32)(end_col - grp_first_col); }
  };
  auto close_tile = [&](i64 end_col) {
    if (cur_cols == 0) return;
    if (grp_cols > 0) { max_slot = max(max_slot, (int)grp_nnz); cur_groups++; close_group(end_col); }
    if (p.emit) {
      const i64 g0 = p.colptr[tile_first_col] - 1;
      TileHdr h;
      h.c0 = (int)tile_first_col; h.ncol = (int)(end_col - tile_first_col); h.nnodes = cur_nodes; h.npairs = (int)cur_pairs;
      h.g0lo = (u32)(g0 & 0xffffffffll); h.g0hi = (u32)(g0 >> 32); h.nnz = (int)cur_nnz; h.blob16 = 0;
      h.pair_base = (u32)tile_pair_base; h.node_base = (u32)(node_base + tile_node_base);
      h.cols_off = p.cols_off;
      h.blob_bytes = blob_size(p.cols_off, (u32)h.ncol, (u32)cur_pairs, (u32)cur_nodes);
      p.hdr[tile_base + cur_tile] = h;
      for (int w2 = cur_groups; w2 <= NW; w2++) cur_grp[w2] = make_uint4((u32)h.ncol, (u32)cur_nnz, (u32)cur_pairs, 0);
      for (int w2 = 0; w2 <= NW; w2++) p.groups[(size_t)(tile_base + cur_tile) * (NW + 1) + w2] = cur_grp[w2];
    }
    cur_groups = 0; grp_cols = 0; grp_nnz = 0;
    tile_node_base = nodes_total;
    cur_tile++; cur_cols = 0; cur_nodes = 0; cur_nnz = 0; cur_pairs = 0;
    if (cur_tile >= (1 << 19)) p.flags[3] = 1;       // generation field of the hash tags
  };
  for (i64 j = j0; j < j1; j++) {
    const i64 kb = p.pairbeg[j], ke = p.pairbeg[j + 1];
    const i64 len = p.colptr[j + 1] - p.colptr[j];
    if (ke == kb || p.col_closed[j] == CK_VERTEX) { close_tile(j); continue; }
    const int n = (int)(ke - kb);
    const u32 P0 = (u32)p.col_P[j], Q0 = (u32)p.col_Q[j];
    for (int attempt = 0; attempt < 2; attempt++) {
      // distinct nodes of the column that the open tile does not hold yet
      u32 fresh_list[2 * MAXRING + 2];
      int fresh = 0;
      auto consider = [&](u32 v) {
        if (lookup(v, false) >= 0) return;
        for (int i = 0; i < fresh; i++) if (fresh_list[i] == v) return;
        fresh_list[fresh++] = v;
      };
      consider(P0); consider(Q0);
      for (int t = 0; t < n; t++) { consider((u32)p.pair_in[kb + t]); consider((u32)p.pair_out[kb + t]); }
      const bool over = (i64)blob_size(p.cols_off, (u32)(cur_cols + 1), (u32)(cur_pairs + n), (u32)(cur_nodes + fresh)) > p.blob_cap;
      const bool grp_full = grp_cols > 0 && (grp_cols == 32 || grp_nnz + len > p.slot_cap);
      if (cur_cols > 0 && ((grp_full && cur_groups + 1 >= NW) || over || cur_nodes + fresh > MAX_TILE_NODES || cur_pairs + n > 65535)) { close_tile(j); continue; }
      if (len > p.slot_cap) { p.flags[3] = 2; return; }
      if (cur_cols == 0) { tile_first_col = j; tile_pair_base = kb; }
      if (grp_full) { max_slot = max(max_slot, (int)grp_nnz); cur_groups++; grp_cols = 0; grp_nnz = 0; close_group(j); }
      if (grp_cols == 0) { cur_grp[cur_groups] = make_uint4((u32)cur_cols, (u32)cur_nnz, (u32)cur_pairs, 0); grp_first_col = j; }
      if (p.emit) p.col_abase[j] = (u32)grp_nnz;
      grp_cols++; grp_nnz += len;
      const int lp = lookup(P0, true), lq = lookup(Q0, true);
      for (int t = 0; t < n; t++) {
        const int li = lookup((u32)p.pair_in[kb + t], true), lo = lookup((u32)p.pair_out[kb + t], true);
        if (p.emit) p.pair_io[kb + t] = (u32)li | ((u32)lo << 12);
      }
      if (p.emit) { p.col_pq[j] = (u32)lp | ((u32)lq << 12); p.col_tile[j] = (u32)(tile_base + cur_tile); }
      cur_cols++; cur_nnz += len; cur_pairs += n;
      break;
    }
  }
  close_tile(j1);
  if (!p.emit) { p.chunk_ntiles[chunk] = cur_tile; p.chunk_nnodes[chunk] = nodes_total; }
  atomicMax(p.flags + 2, max_slot);
}

// per tile: blob size / mirror candidates (inputs of the two scans), then the scanned offsets back into the headers
__global__ void tile_sizes(const TileHdr* hdr, int ntiles, i64* blob16, int* mir) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntiles) return;
  blob16[t] = t < ntiles ? (i64)(hdr[t].blob_bytes / 16) : 0;
  mir[t] = t < ntiles ? 5 * hdr[t].ncol + hdr[t].npairs : 0;
}
__global__ void tile_offsets(TileHdr* hdr, int ntiles, const i64* blob16_scan, uint2* tile_dir, int* maxblob) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntiles) return;
  hdr[t].blob16 = (u32)blob16_scan[t];
  tile_dir[t] = make_uint2((u32)blob16_scan[t], hdr[t].blob_bytes);
  atomicMax(maxblob, (int)hdr[t].blob_bytes);
}
__global__ void flag_vertex_columns(const unsigned char* col_closed, const i64* pairbeg, i64 ncols, i64 ncols_owned, unsigned char* isv) {
  const i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (j < ncols) isv[j] = (j < ncols_owned && col_closed[j] == CK_VERTEX && pairbeg[j + 1] > pairbeg[j]) ? 1 : 0;
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

template <int NW> int set_smem_attr(int bytes) {
  GRMP_CUDA(cudaFuncSetAttribute(p2tet_edge_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return GRMP_OK;
}

__global__ void p2tet_quality_kernel(const double* __restrict__ coords, const i32* __restrict__ cellnodes, i64 ncells, unsigned long long* out) {
  const i64 cell = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  double kap = 0.0;
  if (cell < ncells) {
    const i32* cn = cellnodes + cell * 4;
    double x[4][3];
    for (int a = 0; a < 4; a++) for (int k = 0; k < 3; k++) x[a][k] = coords[(i64)(cn[a] - 1) * 3 + k];
    // inward face normals (unnormalised): n_a is orthogonal to the face opposite to vertex a; S_ab ~ n_a . n_b
    double n[4][3];
    for (int a = 0; a < 4; a++) {
      const int i = (a + 1) & 3, j = (a + 2) & 3, k = (a + 3) & 3;
      const double u[3] = {x[j][0] - x[i][0], x[j][1] - x[i][1], x[j][2] - x[i][2]}, v[3] = {x[k][0] - x[i][0], x[k][1] - x[i][1], x[k][2] - x[i][2]};
      double c[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
      const double w[3] = {x[a][0] - x[i][0], x[a][1] - x[i][1], x[a][2] - x[i][2]};
      const double sgn = (c[0] * w[0] + c[1] * w[1] + c[2] * w[2]) < 0 ? -1.0 : 1.0;
      for (int d = 0; d < 3; d++) n[a][d] = sgn * c[d];
    }
    for (int a = 0; a < 4; a++) {
      double off = 0.0;
      for (int b = 0; b < 4; b++) if (b != a) off += fabs(n[a][0] * n[b][0] + n[a][1] * n[b][1] + n[a][2] * n[b][2]);
      const double dg = n[a][0] * n[a][0] + n[a][1] * n[a][1] + n[a][2] * n[a][2];
      kap = fmax(kap, dg > 0.0 ? off / dg : 1e300);
    }
  }
  for (int d = 16; d > 0; d >>= 1) kap = fmax(kap, __shfl_xor_sync(0xffffffffu, kap, d));
  if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(kap));   // non-negative doubles order like integers
}

}  // namespace

int fast_p2tet_quality(grmp_ctx* ctx, const BlfLocalParams& p, double* kappa) {
  DevBuf<unsigned long long> d;
  GRMP_TRY(d.alloc(1));
  GRMP_CUDA(cudaMemsetAsync(d.p, 0, 8, ctx->stream));
  if (p.g.ncells > 0) {
    p2tet_quality_kernel<<<(unsigned)((p.g.ncells + 255) / 256), 256, 0, ctx->stream>>>(p.g.coords, p.g.cellnodes, p.g.ncells, d.p);
    GRMP_CUDA(cudaGetLastError());
  }
  unsigned long long bits = 0;
  GRMP_CUDA(cudaMemcpyAsync(&bits, d.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  GRMP_CUDA(cudaStreamSynchronize(ctx->stream));
  double k;
  memcpy(&k, &bits, 8);
  *kappa = k;
  return GRMP_OK;
}

namespace {
}  // namespace

bool fast_p2tet_applicable(const BlfLocalParams& p) {
  return p.g.dim == 3 && p.same_eval && p.e1.fam == FAM_H1 && p.e1.op == GRMP_OP_GRAD && p.e1.ncomp == 1 && p.e1.nd == 10 &&
         p.e1.tab_nd == 10 && p.action == GRMP_ACT_NONE && p.reg.n == 0 &&
         (p.apt == GRMP_APT_SYMMETRIC || (p.apt == GRMP_APT_BILINEARFORM && !p.transposed));
}

// Device build: ring orders, tiles, records, mirror lists and vertex-column records without a round trip through the host
// (only counters come back).  Same outputs as the host build below.
static int fast_p2tet_build_device(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, const DofGather& dg, i64 ncols_owned_eff, int NW,
                                   int NBUF, i64 SLOT_CAP, i64 BLOB_CAP, u32 cols_off, FastP2Tet* out) {
  cudaStream_t s = ctx->stream;
  const bool verbose = getenv("GRMP_VERBOSE") != nullptr;
  auto t_start = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!verbose) return;
    cudaStreamSynchronize(s);
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[grmp fast build, device] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t_start).count());
    t_start = t;
  };
  const i64 ncells = p.g.ncells, ncols = pat.ncols, npairs = dg.ncontrib;
  const i64 np1 = std::max<i64>(npairs, 1), nc1 = std::max<i64>(ncols, 1);
  DevBuf<u32> d_cell, d_io, d_code, d_colof, d_coltile, d_colpq, d_abase, d_colgroup, d_colgcount;
  DevBuf<i32> d_in, d_out, d_P, d_Q;
  DevBuf<unsigned char> d_closed;
  DevBuf<int> d_flags;
  GRMP_TRY(d_cell.alloc(np1)); GRMP_TRY(d_io.alloc(np1)); GRMP_TRY(d_code.alloc(np1)); GRMP_TRY(d_colof.alloc(np1));
  GRMP_TRY(d_in.alloc(np1)); GRMP_TRY(d_out.alloc(np1)); GRMP_TRY(d_P.alloc(nc1)); GRMP_TRY(d_Q.alloc(nc1)); GRMP_TRY(d_closed.alloc(nc1));
  GRMP_TRY(d_coltile.alloc(nc1)); GRMP_TRY(d_colpq.alloc(nc1)); GRMP_TRY(d_abase.alloc(nc1)); GRMP_TRY(d_colgroup.alloc(nc1)); GRMP_TRY(d_colgcount.alloc(nc1));
  GRMP_TRY(d_flags.alloc(4));
  GRMP_CUDA(cudaMemsetAsync(d_flags.p, 0, 16, s));
  GRMP_CUDA(cudaMemsetAsync(d_io.p, 0, d_io.bytes(), s)); GRMP_CUDA(cudaMemsetAsync(d_code.p, 0, d_code.bytes(), s));
  GRMP_CUDA(cudaMemsetAsync(d_coltile.p, 0xff, d_coltile.bytes(), s)); GRMP_CUDA(cudaMemsetAsync(d_colpq.p, 0, d_colpq.bytes(), s));
  GRMP_CUDA(cudaMemsetAsync(d_abase.p, 0, d_abase.bytes(), s)); GRMP_CUDA(cudaMemsetAsync(d_colgroup.p, 0, d_colgroup.bytes(), s));
  GRMP_CUDA(cudaMemsetAsync(d_colgcount.p, 0, d_colgcount.bytes(), s));
  GRMP_CUDA(cudaMemsetAsync(d_P.p, 0, d_P.bytes(), s)); GRMP_CUDA(cudaMemsetAsync(d_Q.p, 0, d_Q.bytes(), s));
  // (2a) ring order of every edge column
  RingParams rp{dg.gcell.p, dg.gsrc.p, dg.segptr.p, pat.colptr.p, p.g.cellnodes, ncells, ncols, ncols_owned_eff,
                d_cell.p, d_code.p, d_colof.p, d_in.p, d_out.p, d_P.p, d_Q.p, d_closed.p, d_flags.p};
  if (ncols) ring_order_kernel<<<(unsigned)((ncols + 127) / 128), 128, 0, s>>>(rp);
  GRMP_CUDA(cudaGetLastError());
  int hflags[4] = {0, 0, 0, 0};
  GRMP_CUDA(cudaMemcpyAsync(hflags, d_flags.p, 16, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  if (hflags[0]) {
    static const char* msg[] = {"", "fast path: an edge column has more than 254 entries", "fast path: mixed dof types in one column",
                                "fast path: inconsistent edge column", "fast path: non-manifold edge star", "fast path: edge star is not a single ring",
                                "fast path: edge star mixes a ring and chains", "fast path: more than 64 cells around an edge (GRMP_FAST_HOST_BUILD=1 handles 255)",
                                "fast path: an owned edge star is not a single chain", "fast path: edge ring does not close"};
    return fail(GRMP_EUNSUPPORTED, msg[std::min(hflags[0], 9)]);
  }
  const bool any_end = hflags[1] != 0;
  lap("ring order");
  // (2b) tiles: chunks of TB_CHUNK columns packed independently, count pass + emit pass
  const i64 nchunks = (ncols + TB_CHUNK - 1) / TB_CHUNK;
  DevBuf<int> d_cnt_t, d_cnt_n, d_off_t, d_off_n;
  GRMP_TRY(d_cnt_t.alloc(nchunks + 1)); GRMP_TRY(d_cnt_n.alloc(nchunks + 1)); GRMP_TRY(d_off_t.alloc(nchunks + 1)); GRMP_TRY(d_off_n.alloc(nchunks + 1));
  GRMP_CUDA(cudaMemsetAsync(d_cnt_t.p, 0, d_cnt_t.bytes(), s)); GRMP_CUDA(cudaMemsetAsync(d_cnt_n.p, 0, d_cnt_n.bytes(), s));
  TileBuildParams tb{};
  tb.pairbeg = dg.segptr.p; tb.colptr = pat.colptr.p; tb.col_closed = d_closed.p; tb.pair_in = d_in.p; tb.pair_out = d_out.p; tb.col_P = d_P.p; tb.col_Q = d_Q.p;
  tb.ncols = ncols; tb.nw = NW; tb.slot_cap = SLOT_CAP; tb.blob_cap = BLOB_CAP; tb.cols_off = cols_off;
  tb.chunk_ntiles = d_cnt_t.p; tb.chunk_nnodes = d_cnt_n.p; tb.emit = 0; tb.flags = d_flags.p;
  if (nchunks) tile_build_kernel<<<(unsigned)nchunks, 32, 0, s>>>(tb);
  GRMP_CUDA(cudaGetLastError());
  DevBuf<unsigned char> d_tmp;
  auto scan_int = [&](const int* in, int* outp, i64 n) -> int {
    size_t tbytes = 0;
    GRMP_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tbytes, in, outp, n, s));
    if (tbytes > d_tmp.n) GRMP_TRY(d_tmp.alloc(std::max<size_t>(tbytes, 16)));
    GRMP_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, tbytes, in, outp, n, s));
    return GRMP_OK;
  };
  GRMP_TRY(scan_int(d_cnt_t.p, d_off_t.p, nchunks + 1));
  GRMP_TRY(scan_int(d_cnt_n.p, d_off_n.p, nchunks + 1));
  int tot[2] = {0, 0};
  GRMP_CUDA(cudaMemcpyAsync(&tot[0], d_off_t.p + nchunks, 4, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaMemcpyAsync(&tot[1], d_off_n.p + nchunks, 4, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaMemcpyAsync(hflags, d_flags.p, 16, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  if (hflags[3] == 2) return fail(GRMP_EUNSUPPORTED, "fast path: a single column exceeds the stage slot");
  if (hflags[3]) return fail(GRMP_EUNSUPPORTED, "fast path: too many tiles in one chunk");
  const int ntiles = tot[0];
  const i64 nnodes_total = tot[1];
  DevBuf<uint4> d_groups;
  DevBuf<u32>& d_nodeids = out->tile_nodeids;
  DevBuf<int4>& d_hdr = out->tile_hdr;
  GRMP_TRY(d_hdr.alloc((size_t)std::max(ntiles, 1) * 3)); GRMP_TRY(d_groups.alloc((size_t)std::max(ntiles, 1) * (NW + 1)));
  GRMP_TRY(d_nodeids.alloc(std::max<i64>(nnodes_total, 1)));
  GRMP_CUDA(cudaMemsetAsync(d_hdr.p, 0, d_hdr.bytes(), s)); GRMP_CUDA(cudaMemsetAsync(d_groups.p, 0, d_groups.bytes(), s));
  GRMP_CUDA(cudaMemsetAsync(d_nodeids.p, 0, d_nodeids.bytes(), s));
  tb.chunk_ntiles = d_off_t.p; tb.chunk_nnodes = d_off_n.p; tb.emit = 1;
  tb.hdr = reinterpret_cast<TileHdr*>(d_hdr.p); tb.groups = d_groups.p; tb.tile_nodeids = d_nodeids.p; tb.col_tile = d_coltile.p; tb.col_pq = d_colpq.p;
  tb.col_abase = d_abase.p; tb.col_group = d_colgroup.p; tb.col_gcount = d_colgcount.p; tb.pair_io = d_io.p;
  if (nchunks) tile_build_kernel<<<(unsigned)nchunks, 32, 0, s>>>(tb);
  GRMP_CUDA(cudaGetLastError());
  // blob offsets and mirror-candidate offsets: scans over the tiles
  DevBuf<i64> d_b16, d_b16s;
  DevBuf<int> d_mir, d_mirbase, d_maxblob;
  GRMP_TRY(d_b16.alloc(ntiles + 1)); GRMP_TRY(d_b16s.alloc(ntiles + 1)); GRMP_TRY(d_mir.alloc(ntiles + 1)); GRMP_TRY(d_mirbase.alloc(ntiles + 1));
  GRMP_TRY(d_maxblob.alloc(1));
  GRMP_CUDA(cudaMemsetAsync(d_maxblob.p, 0, 4, s));
  tile_sizes<<<(unsigned)((ntiles + 1 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const TileHdr*>(d_hdr.p), ntiles, d_b16.p, d_mir.p);
  {
    size_t tbytes = 0;
    GRMP_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tbytes, d_b16.p, d_b16s.p, ntiles + 1, s));
    if (tbytes > d_tmp.n) GRMP_TRY(d_tmp.alloc(std::max<size_t>(tbytes, 16)));
    GRMP_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp.p, tbytes, d_b16.p, d_b16s.p, ntiles + 1, s));
  }
  GRMP_TRY(scan_int(d_mir.p, d_mirbase.p, ntiles + 1));
  GRMP_TRY(out->tile_dir.alloc(std::max(ntiles, 1)));
  GRMP_CUDA(cudaMemsetAsync(out->tile_dir.p, 0, out->tile_dir.bytes(), s));
  if (ntiles) tile_offsets<<<(unsigned)((ntiles + 255) / 256), 256, 0, s>>>(reinterpret_cast<TileHdr*>(d_hdr.p), ntiles, d_b16s.p, out->tile_dir.p, d_maxblob.p);
  GRMP_CUDA(cudaGetLastError());
  i64 blob_total16 = 0;
  int nmir_total = 0, max_blob = 0;
  GRMP_CUDA(cudaMemcpyAsync(&blob_total16, d_b16s.p + ntiles, 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaMemcpyAsync(&nmir_total, d_mirbase.p + ntiles, 4, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaMemcpyAsync(&max_blob, d_maxblob.p, 4, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaMemcpyAsync(hflags, d_flags.p, 16, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  const i64 max_slot = hflags[2];
  lap("tiles (chunks)");
  const i64 slot_elems = (max_slot + 2 + 1) & ~1ll;
  const i64 max_smem = (i64)NBUF * max_blob + NW * 8 * slot_elems;
  if (max_smem > 220 * 1024) return fail(GRMP_EUNSUPPORTED, "fast path: a single column exceeds the shared-memory tile");
  if (max_blob > 8 * 65535) return fail(GRMP_EUNSUPPORTED, "fast path: tile blob exceeds the 16-bit word index of the mirror list");
  if (blob_total16 >= (i64)NONE) return fail(GRMP_EUNSUPPORTED, "fast path: record blob exceeds 64 GB");
  out->ntiles = ntiles; out->npairs = npairs; out->smem_bytes = (int)std::max<i64>(max_smem, 1024); out->in_stride = (u32)max_blob;
  out->slot_elems = (u32)slot_elems;
  out->prof.release();
  if (getenv("GRMP_FAST_PROF")) {
    GRMP_TRY(out->prof.alloc(8 * 1024));
    GRMP_CUDA(cudaMemsetAsync(out->prof.p, 0, out->prof.bytes(), s));
  }
  GRMP_TRY(out->tile_counter.alloc(1));
  GRMP_CUDA(cudaMemsetAsync(out->tile_counter.p, 0, sizeof(unsigned long long), s));
  // (3) records into the blobs, mirror candidates sorted by destination
  DevBuf<u32> d_mkey, d_mval, d_mkey2, d_mval2;
  const i64 nm1 = std::max<i64>(nmir_total, 1);
  GRMP_TRY(d_mkey.alloc(nm1)); GRMP_TRY(d_mval.alloc(nm1)); GRMP_TRY(d_mkey2.alloc(nm1)); GRMP_TRY(d_mval2.alloc(nm1));
  GRMP_CUDA(cudaMemsetAsync(d_mkey.p, 0xff, d_mkey.bytes(), s));
  GRMP_TRY(out->blob.alloc(std::max<size_t>((size_t)blob_total16 * 16, 16)));
  GRMP_CUDA(cudaMemsetAsync(out->blob.p, 0, out->blob.bytes(), s));
  out->end_slots.release();
  if (any_end) GRMP_TRY(out->end_slots.alloc(np1));
  PackParams pp{d_cell.p, d_io.p, d_code.p, dg.segptr.p, pat.colptr.p, pat.rowval.p, p.e1.celldofs, d_colof.p, d_closed.p,
                d_coltile.p, d_colpq.p, d_abase.p, d_colgroup.p, d_colgcount.p, reinterpret_cast<const TileHdr*>(d_hdr.p), d_mirbase.p, npairs, ncols,
                out->blob.p, out->end_slots.p, (u32)std::min<i64>(out->halo_first, (i64)NONE), d_mkey.p, d_mval.p};
  if (npairs && ntiles) pack_pairs<<<(unsigned)((npairs + 255) / 256), 256, 0, s>>>(pp);
  if (ncols && ntiles) pack_cols<<<(unsigned)((ncols + 255) / 256), 256, 0, s>>>(pp);
  GRMP_CUDA(cudaGetLastError());
  if (ntiles > 0) {
    size_t tmp_bytes = 0;
    GRMP_CUDA(cub::DeviceSegmentedSort::SortPairs(nullptr, tmp_bytes, d_mkey.p, d_mkey2.p, d_mval.p, d_mval2.p, nmir_total, ntiles,
                                                  d_mirbase.p, d_mirbase.p + 1, s));
    DevBuf<unsigned char> d_tmp2;
    GRMP_TRY(d_tmp2.alloc(std::max<size_t>(tmp_bytes, 16)));
    GRMP_CUDA(cub::DeviceSegmentedSort::SortPairs(d_tmp2.p, tmp_bytes, d_mkey.p, d_mkey2.p, d_mval.p, d_mval2.p, nmir_total, ntiles,
                                                  d_mirbase.p, d_mirbase.p + 1, s));
    pack_tile_rest<<<ntiles, 128, 0, s>>>(reinterpret_cast<const TileHdr*>(d_hdr.p), d_groups.p, NW, d_nodeids.p, p.g.coords, d_mirbase.p,
                                          d_mkey2.p, d_mval2.p, out->blob.p, out->tile_dir.p);
    GRMP_CUDA(cudaGetLastError());
    GRMP_CUDA(cudaStreamSynchronize(s));
  }
  lap("pack kernels + mirror sort");
  // (4) vertex columns: compacted list + diagonal slots
  {
    DevBuf<unsigned char> d_isv;
    DevBuf<i64> d_nsel;
    GRMP_TRY(d_isv.alloc(nc1)); GRMP_TRY(d_nsel.alloc(1)); GRMP_TRY(out->vcols.alloc(nc1));
    if (ncols) flag_vertex_columns<<<(unsigned)((ncols + 255) / 256), 256, 0, s>>>(d_closed.p, dg.segptr.p, ncols, ncols_owned_eff, d_isv.p);
    GRMP_CUDA(cudaMemsetAsync(d_nsel.p, 0, 8, s));
    if (ncols) {
      cub::CountingInputIterator<u32> it(0);
      size_t tbytes = 0;
      GRMP_CUDA(cub::DeviceSelect::Flagged(nullptr, tbytes, it, d_isv.p, out->vcols.p, d_nsel.p, (int)ncols, s));
      if (tbytes > d_tmp.n) GRMP_TRY(d_tmp.alloc(std::max<size_t>(tbytes, 16)));
      GRMP_CUDA(cub::DeviceSelect::Flagged(d_tmp.p, tbytes, it, d_isv.p, out->vcols.p, d_nsel.p, (int)ncols, s));
    }
    i64 nsel = 0;
    GRMP_CUDA(cudaMemcpyAsync(&nsel, d_nsel.p, 8, cudaMemcpyDeviceToHost, s));
    GRMP_CUDA(cudaStreamSynchronize(s));
    out->nvcols = nsel;
    GRMP_TRY(out->vrec.alloc((size_t)std::max<i64>(2 * nsel, 2)));
    if (nsel > 0) {
      find_diag_slots<<<(unsigned)((nsel + 255) / 256), 256, 0, s>>>(out->vcols.p, nsel, pat.colptr.p, pat.rowval.p, d_closed.p, out->vrec.p);
      GRMP_CUDA(cudaGetLastError());
    }
  }
  const int smem_attr = (int)std::max<i64>(max_smem, 1024);
  GRMP_TRY(set_smem_attr<3>(smem_attr)); GRMP_TRY(set_smem_attr<4>(smem_attr)); GRMP_TRY(set_smem_attr<5>(smem_attr));
  GRMP_TRY(set_smem_attr<6>(smem_attr)); GRMP_TRY(set_smem_attr<7>(smem_attr));
  GRMP_CUDA(cudaStreamSynchronize(s));
  lap("vertex columns");
  return GRMP_OK;
}

int fast_p2tet_build(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, const std::vector<double>& w,
                     const std::vector<double>& derivs, i64 ncols_owned, i64 geom_version, FastP2Tet* out) {
  // halo columns (>= ncols_owned) are processed too: their mirrors complete the owned vertex columns;
  // only they may consist of several chains (cells around a halo edge are present only where they touch an owned dof)
  cudaStream_t s = ctx->stream;
  const bool verbose = getenv("GRMP_VERBOSE") != nullptr;
  auto t_start = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!verbose) return;
    cudaStreamSynchronize(s);
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[grmp fast build] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t_start).count());
    t_start = t;
  };
  const i64 ncells = p.g.ncells, ncols = pat.ncols, nnodes = p.g.nnodes;
  const i64 ncols_owned_eff = (ncols_owned >= 0 && ncols_owned < ncols) ? ncols_owned : ncols;
  out->ntiles = 0; out->nvcols = 0;
  out->halo_first = pat.nnz;
  if (ncols_owned_eff < ncols) {
    i64 cpv = 0;
    GRMP_CUDA(cudaMemcpyAsync(&cpv, pat.colptr.p + ncols_owned_eff, 8, cudaMemcpyDeviceToHost, s));
    GRMP_CUDA(cudaStreamSynchronize(s));
    out->halo_first = cpv - 1;
  }
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
  if (npairs >= (i64)NONE) return fail(GRMP_EUNSUPPORTED, "fast path: more than 2^32-1 pairs");
  // tile shape (tunable for experiments: GRMP_FAST_NW in 3..7, GRMP_FAST_SLOT, GRMP_FAST_SMEM_KB)
  int NW = getenv("GRMP_FAST_NW") ? atoi(getenv("GRMP_FAST_NW")) : NW_DEFAULT;
  if (NW < 3 || NW > 7) NW = NW_DEFAULT;
  const i64 SLOT_CAP = getenv("GRMP_FAST_SLOT") ? std::max(256, atoi(getenv("GRMP_FAST_SLOT"))) : SLOT_DEFAULT;   // nzval entries per group
  const i64 SMEM_BUDGET = getenv("GRMP_FAST_SMEM_KB") ? 1024 * (i64)atoi(getenv("GRMP_FAST_SMEM_KB")) : smem_budget_default(NW);
  // shared memory of a CTA: 2 input buffers (largest blob) + NW stage slots
  int NBUF = getenv("GRMP_FAST_NBUF") ? atoi(getenv("GRMP_FAST_NBUF")) : 2;
  if (NBUF != 2 && NBUF != 3) NBUF = 2;
  const i64 BLOB_CAP = ((SMEM_BUDGET - NW * 8 * (SLOT_CAP + 2)) / NBUF) & ~15ll;
  if (BLOB_CAP < 4096) return fail(GRMP_EUNSUPPORTED, "fast path: shared-memory budget too small for the tile shape");
  const u32 cols_off = 48u + 16u * (u32)(NW + 1);
  out->nw = NW;
  out->nbuf = NBUF;
  out->geom_version = geom_version;
  const bool device_build = !getenv("GRMP_FAST_HOST_BUILD") && BLOB_CAP / 24 < TB_HASH / 2;
  if (device_build) return fast_p2tet_build_device(ctx, p, pat, dg, ncols_owned_eff, NW, NBUF, SLOT_CAP, BLOB_CAP, cols_off, out);
  std::vector<u32> h_cell(npairs), h_src(npairs);
  std::vector<i64> h_pairbeg(ncols + 1), h_colptr(ncols + 1);
  std::vector<i32> h_cn((size_t)ncells * 4);
  GRMP_CUDA(cudaMemcpyAsync(h_cell.data(), dg.gcell.p, npairs * 4, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaMemcpyAsync(h_src.data(), dg.gsrc.p, npairs * 4, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaMemcpyAsync(h_pairbeg.data(), dg.segptr.p, (ncols + 1) * 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaMemcpyAsync(h_colptr.data(), pat.colptr.p, (ncols + 1) * 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaMemcpyAsync(h_cn.data(), p.g.cellnodes, (size_t)ncells * 16, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  lap("dof gather + downloads");
  // (2) host: ring order of every edge column, tiles over the edge columns, list of vertex columns
  std::vector<u32> pair_cell(npairs), pair_io(npairs), pair_code(npairs), col_of_pair(npairs), vcols;
  std::vector<unsigned char> col_closed(ncols, 2);
  std::vector<u32> col_tile(ncols, NONE), col_pq(ncols, 0), col_abase(ncols, 0);
  std::vector<TileHdr> hdr;
  std::vector<u32> tile_nodeids;
  std::vector<uint4> groups;                    // (NW+1) per tile: first column (tile-local), first slot (relative to g0), first pair (tile-local)
  std::vector<u32> col_group(ncols, 0), col_gcount(ncols, 0);   // first column / #columns of every edge column's group
  i64 grp_first_col = 0;
  std::vector<int> mir_base(1, 0);              // first mirror candidate of every tile
  std::vector<i32> nmark(nnodes + 1, -1), nlocal(nnodes + 1, 0);
  int cur_tile = 0, cur_cols = 0, cur_nodes = 0;
  int cur_groups = 0, grp_cols = 0;             // closed groups of the open tile; columns / slots of the open group
  i64 grp_nnz = 0;
  uint4 cur_grp[9];
  auto close_group = [&](i64 end_col) {       // columns [grp_first_col, end_col) form a group
    for (i64 c = grp_first_col; c < end_col; c++) { col_group[c] = (u32)grp_first_col; col_gcount[c] = (u32)(end_col - grp_first_col); }
  };
  i64 cur_nnz = 0, cur_pairs = 0, tile_first_col = 0, tile_node_base = 0, tile_pair_base = 0;
  i64 max_blob = 0, max_slot = 0, blob_total16 = 0;
  bool any_end = false;
  auto close_tile = [&](i64 end_col) {
    if (cur_cols == 0) return;
    const i64 g0 = h_colptr[tile_first_col] - 1;
    TileHdr h;
    h.c0 = (int)tile_first_col; h.ncol = (int)(end_col - tile_first_col); h.nnodes = cur_nodes; h.npairs = (int)cur_pairs;
    h.g0lo = (u32)(g0 & 0xffffffffll); h.g0hi = (u32)(g0 >> 32); h.nnz = (int)cur_nnz; h.blob16 = (u32)blob_total16;
    h.pair_base = (u32)tile_pair_base; h.node_base = (u32)tile_node_base;
    h.cols_off = cols_off;
    h.blob_bytes = blob_size(cols_off, (u32)h.ncol, (u32)cur_pairs, (u32)cur_nodes);
    hdr.push_back(h);
    blob_total16 += h.blob_bytes / 16;
    max_blob = std::max<i64>(max_blob, h.blob_bytes);
    mir_base.push_back(mir_base.back() + 5 * h.ncol + (int)cur_pairs);
    if (grp_cols > 0) { max_slot = std::max<i64>(max_slot, grp_nnz); cur_groups++; close_group(end_col); }
    for (int w2 = cur_groups; w2 <= NW; w2++) cur_grp[w2] = make_uint4((u32)h.ncol, (u32)cur_nnz, (u32)cur_pairs, 0);   // empty trailing groups
    for (int w2 = 0; w2 <= NW; w2++) groups.push_back(cur_grp[w2]);
    cur_groups = 0; grp_cols = 0; grp_nnz = 0;
    tile_node_base = (i64)tile_nodeids.size();
    cur_tile++; cur_cols = 0; cur_nodes = 0; cur_nnz = 0; cur_pairs = 0;
  };
  // (2a) ring order of every edge column: independent per column -> host threads.  Stores the ring as global node ids
  //      (pair_in / pair_out), the orientation (col_P, col_Q) and the flags; errors are reported after the join.
  std::vector<i32> pair_in(npairs), pair_out(npairs), col_P(ncols, 0), col_Q(ncols, 0);
  {
    unsigned nthr = std::thread::hardware_concurrency();
    nthr = std::max(1u, std::min(nthr ? nthr : 1u, 32u));
    if (getenv("GRMP_HOST_THREADS")) nthr = (unsigned)std::max(1, atoi(getenv("GRMP_HOST_THREADS")));
    if (ncols < 4096) nthr = 1;
    std::vector<const char*> err(nthr, nullptr);
    std::vector<char> end_seen(nthr, 0);
    auto work = [&](unsigned tix) {
      struct RP { u32 cell; int P, Q, R, S; i32 nR, nS; };
      std::vector<RP> rp;
      std::vector<char> used;
      std::vector<int> degR, degS;
      const i64 j0 = ncols * tix / nthr, j1 = ncols * (tix + 1) / nthr;
      for (i64 j = j0; j < j1; j++) {
        const i64 kb = h_pairbeg[j], ke = h_pairbeg[j + 1];
        const i64 len = h_colptr[j + 1] - h_colptr[j];
        if (ke == kb) continue;
        const int lj0 = (int)(h_src[kb] / (u32)ncells);
        if (lj0 < 4) {   // vertex column: filled by mirrors + the diagonal kernel
          for (i64 k = kb; k < ke; k++) { pair_cell[k] = h_cell[k]; pair_io[k] = 0; pair_code[k] = 0; col_of_pair[k] = (u32)j; }
          continue;
        }
        if (len > 254 || ke - kb > 255) { err[tix] = "fast path: an edge column has more than 254 entries"; return; }
        // ---- ring order of the cells around the edge ----
        const int n = (int)(ke - kb);
        rp.resize(n);
        i32 P0 = 0, Q0 = 0;
        for (int t = 0; t < n; t++) {
          const u32 c = h_cell[kb + t];
          const int lj = (int)(h_src[kb + t] / (u32)ncells);
          if (lj < 4) { err[tix] = "fast path: mixed dof types in one column"; return; }
          int pl, ql; edge_nodes(lj - 4, pl, ql);
          int rl = -1, sl = -1;
          for (int v = 0; v < 4; v++) if (v != pl && v != ql) { if (rl < 0) rl = v; else sl = v; }
          const i32* cn = &h_cn[(size_t)c * 4];
          if (t == 0) { P0 = cn[pl]; Q0 = cn[ql]; }
          if (cn[pl] != P0) { int tmp = pl; pl = ql; ql = tmp; }      // consistent global orientation (P,Q)
          if (cn[pl] != P0 || cn[ql] != Q0) { err[tix] = "fast path: inconsistent edge column"; return; }
          rp[t] = RP{c, pl, ql, rl, sl, cn[rl], cn[sl]};
        }
        col_P[j] = P0; col_Q[j] = Q0;
        // degrees of the ring vertices; chains start at vertices of degree 1, a star without such a vertex is a closed ring
        degR.resize(n); degS.resize(n);
        bool closed = true;
        for (int t = 0; t < n; t++) {
          int dR = 0, dS = 0;
          for (int u = 0; u < n; u++) {
            dR += (rp[u].nR == rp[t].nR) + (rp[u].nS == rp[t].nR);
            dS += (rp[u].nR == rp[t].nS) + (rp[u].nS == rp[t].nS);
          }
          if (dR > 2 || dS > 2) { err[tix] = "fast path: non-manifold edge star"; return; }
          degR[t] = dR; degS[t] = dS;
          if (dR == 1 || dS == 1) closed = false;
        }
        used.assign(n, 0);
        int step = 0, nchains = 0;
        while (step < n) {
          int start = -1, start_in_is_R = 1;
          if (closed) { if (step != 0) { err[tix] = "fast path: edge star is not a single ring"; return; } start = 0; }
          else
            for (int t = 0; t < n && start < 0; t++)
              if (!used[t]) { if (degR[t] == 1) { start = t; start_in_is_R = 1; } else if (degS[t] == 1) { start = t; start_in_is_R = 0; } }
          if (start < 0) { err[tix] = "fast path: edge star mixes a ring and chains"; return; }
          if (nchains > 0 && j < ncols_owned_eff) { err[tix] = "fast path: an owned edge star is not a single chain"; return; }
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
            u32 fl = 0u;
            if (chain_first && nchains > 0) fl |= PF_RESET;
            pair_cell[k] = rp[curp].cell;
            pair_code[k] = (u32)(rp[curp].P | (rp[curp].Q << 2) | (I << 4) | (O << 6)) | (fl << 8);
            col_of_pair[k] = (u32)j;
            pair_in[k] = vin; pair_out[k] = vout;
            step++; chain_first = false;
            int nxt = -1;
            for (int u = 0; u < n; u++) if (!used[u] && (rp[u].nR == vout || rp[u].nS == vout)) { nxt = u; break; }
            if (nxt < 0) {
              if (closed && (step != n || vout != first_in)) { err[tix] = "fast path: edge ring does not close"; return; }
              if (!closed && step < n) { pair_code[k] |= (PF_END << 8); end_seen[tix] = 1; }     // a further chain follows
              break;
            }
            vin = vout; curp = nxt;
          }
          nchains++;
        }
        col_closed[j] = closed ? 1 : 0;
      }
    };
    if (nthr == 1) work(0);
    else {
      std::vector<std::thread> pool;
      for (unsigned t = 0; t < nthr; t++) pool.emplace_back(work, t);
      for (auto& th : pool) th.join();
    }
    for (unsigned t = 0; t < nthr; t++) {
      if (err[t]) return fail(GRMP_EUNSUPPORTED, err[t]);
      any_end = any_end || end_seen[t];
    }
  }
  lap("host ring order (threads)");
  // (2b) tiles and warp groups over the edge columns, list of vertex columns: sequential
  std::vector<i32> colnodes;
  std::vector<i64> cmark(nnodes + 1, -1);       // stamp: node already seen in this column (no sort / unique per column)
  for (i64 j = 0; j < ncols; j++) {
    const i64 kb = h_pairbeg[j], ke = h_pairbeg[j + 1];
    const i64 len = h_colptr[j + 1] - h_colptr[j];
    if (ke == kb) { close_tile(j); continue; }
    if (col_closed[j] == 2) {   // vertex column
      close_tile(j);
      if (j < ncols_owned_eff) vcols.push_back((u32)j);
      continue;
    }
    const int n = (int)(ke - kb);
    const i32 P0 = col_P[j], Q0 = col_Q[j];
    const i32* ring_in = &pair_in[kb];
    const i32* ring_out = &pair_out[kb];
    colnodes.clear();
    colnodes.push_back(P0); colnodes.push_back(Q0);
    for (int t = 0; t < n; t++) { colnodes.push_back(ring_in[t]); colnodes.push_back(ring_out[t]); }
    {
      size_t m = 0;
      for (i32 v : colnodes) if (cmark[v] != j) { cmark[v] = j; colnodes[m++] = v; }
      colnodes.resize(m);
    }
    // ---- tile budget ----
    for (int attempt = 0; attempt < 2; attempt++) {
      int fresh = 0;
      for (i32 v : colnodes) if (nmark[v] != cur_tile) fresh++;
      const bool over = (i64)blob_size(cols_off, (u32)(cur_cols + 1), (u32)(cur_pairs + n), (u32)(cur_nodes + fresh)) > BLOB_CAP;
      const bool grp_full = grp_cols > 0 && (grp_cols == 32 || grp_nnz + len > SLOT_CAP);   // the column would open a new group
      if (cur_cols > 0 && ((grp_full && cur_groups + 1 >= NW) || over || cur_nodes + fresh > MAX_TILE_NODES || cur_pairs + n > 65535)) { close_tile(j); continue; }
      if (len > SLOT_CAP) return fail(GRMP_EUNSUPPORTED, "fast path: a single column exceeds the stage slot");
      if (cur_cols == 0) { tile_first_col = j; tile_pair_base = kb; }
      if (grp_full) { max_slot = std::max<i64>(max_slot, grp_nnz); cur_groups++; grp_cols = 0; grp_nnz = 0; close_group(j); }
      if (grp_cols == 0) { cur_grp[cur_groups] = make_uint4((u32)cur_cols, (u32)cur_nnz, (u32)cur_pairs, 0); grp_first_col = j; }
      col_abase[j] = (u32)grp_nnz;
      grp_cols++; grp_nnz += len;
      for (i32 v : colnodes) if (nmark[v] != cur_tile) { nmark[v] = cur_tile; nlocal[v] = cur_nodes++; tile_nodeids.push_back((u32)v); }
      for (int t = 0; t < n; t++) pair_io[kb + t] = (u32)nlocal[ring_in[t]] | ((u32)nlocal[ring_out[t]] << 12);
      col_pq[j] = (u32)nlocal[P0] | ((u32)nlocal[Q0] << 12);
      col_tile[j] = (u32)cur_tile;
      cur_cols++; cur_nnz += len; cur_pairs += n;
      break;
    }
  }
  close_tile(ncols);
  lap("host tiles (sequential)");
  const i64 slot_elems = (max_slot + 2 + 1) & ~1ll;     // even: keeps every slot 16-byte aligned
  const i64 max_smem = NBUF * max_blob + NW * 8 * slot_elems;
  if (max_smem > 220 * 1024) return fail(GRMP_EUNSUPPORTED, "fast path: a single column exceeds the shared-memory tile");
  if (max_blob > 8 * 65535) return fail(GRMP_EUNSUPPORTED, "fast path: tile blob exceeds the 16-bit word index of the mirror list");
  const int ntiles = (int)hdr.size();
  if (blob_total16 >= (i64)NONE) return fail(GRMP_EUNSUPPORTED, "fast path: record blob exceeds 64 GB");
  out->ntiles = ntiles; out->npairs = npairs; out->smem_bytes = (int)std::max<i64>(max_smem, 1024); out->in_stride = (u32)max_blob;
  out->slot_elems = (u32)slot_elems; out->nvcols = (i64)vcols.size();
  if (hdr.empty()) { TileHdr z{}; hdr.push_back(z); groups.assign(NW + 1, make_uint4(0, 0, 0, 0)); mir_base.push_back(0); }
  if (tile_nodeids.empty()) tile_nodeids.push_back(1);
  if (vcols.empty()) vcols.push_back(0);
  std::vector<uint2> tile_dir(hdr.size());
  for (size_t t2 = 0; t2 < hdr.size(); t2++) tile_dir[t2] = make_uint2(hdr[t2].blob16, hdr[t2].blob_bytes);
  DevBuf<uint4> d_groups;
  DevBuf<u32>& d_nodeids = out->tile_nodeids;
  DevBuf<int4>& d_hdr = out->tile_hdr;
  DevBuf<int> d_mirbase;
  GRMP_TRY(d_groups.upload(groups.data(), groups.size(), s));
  GRMP_TRY(d_hdr.upload(reinterpret_cast<const int4*>(hdr.data()), hdr.size() * 3, s));
  GRMP_TRY(d_mirbase.upload(mir_base.data(), mir_base.size(), s));
  GRMP_TRY(out->tile_dir.upload(tile_dir.data(), tile_dir.size(), s));
  GRMP_TRY(d_nodeids.upload(tile_nodeids.data(), tile_nodeids.size(), s));
  out->prof.release();
  if (getenv("GRMP_FAST_PROF")) {
    GRMP_TRY(out->prof.alloc(8 * 1024));
    GRMP_CUDA(cudaMemsetAsync(out->prof.p, 0, out->prof.bytes(), s));
  }
  GRMP_TRY(out->tile_counter.alloc(1));
  GRMP_CUDA(cudaMemsetAsync(out->tile_counter.p, 0, sizeof(unsigned long long), s));
  lap("tile tables upload");
  // (3) pack column / pair records into the tile blobs on the device (slots are looked up in the pattern by (row, col)),
  //     sort every tile's mirror candidates by destination slot
  const i64 nmir_total = mir_base.back();
  DevBuf<u32> d_cell, d_io, d_code, d_colof, d_coltile, d_colpq, d_abase, d_mkey, d_mval, d_mkey2, d_mval2, d_colgroup, d_colgcount;
  DevBuf<unsigned char> d_closed;
  GRMP_TRY(d_cell.upload(pair_cell.data(), npairs, s)); GRMP_TRY(d_io.upload(pair_io.data(), npairs, s));
  GRMP_TRY(d_code.upload(pair_code.data(), npairs, s)); GRMP_TRY(d_colof.upload(col_of_pair.data(), npairs, s));
  GRMP_TRY(d_closed.upload(col_closed.data(), ncols, s));
  GRMP_TRY(d_coltile.upload(col_tile.data(), ncols, s)); GRMP_TRY(d_colpq.upload(col_pq.data(), ncols, s));
  GRMP_TRY(d_abase.upload(col_abase.data(), ncols, s));
  GRMP_TRY(d_colgroup.upload(col_group.data(), ncols, s)); GRMP_TRY(d_colgcount.upload(col_gcount.data(), ncols, s));
  GRMP_TRY(d_mkey.alloc(std::max<i64>(nmir_total, 1))); GRMP_TRY(d_mval.alloc(std::max<i64>(nmir_total, 1)));
  GRMP_TRY(d_mkey2.alloc(std::max<i64>(nmir_total, 1))); GRMP_TRY(d_mval2.alloc(std::max<i64>(nmir_total, 1)));
  GRMP_CUDA(cudaMemsetAsync(d_mkey.p, 0xff, d_mkey.bytes(), s));
  GRMP_TRY(out->blob.alloc(std::max<size_t>((size_t)blob_total16 * 16, 16)));
  GRMP_CUDA(cudaMemsetAsync(out->blob.p, 0, out->blob.bytes(), s));
  out->end_slots.release();
  if (any_end) GRMP_TRY(out->end_slots.alloc(std::max<i64>(npairs, 1)));
  PackParams pp{d_cell.p, d_io.p, d_code.p, dg.segptr.p, pat.colptr.p, pat.rowval.p, p.e1.celldofs, d_colof.p, d_closed.p,
                d_coltile.p, d_colpq.p, d_abase.p, d_colgroup.p, d_colgcount.p, reinterpret_cast<const TileHdr*>(d_hdr.p), d_mirbase.p, npairs, ncols,
                out->blob.p, out->end_slots.p, (u32)std::min<i64>(out->halo_first, (i64)NONE), d_mkey.p, d_mval.p};
  lap("pair arrays upload");
  if (npairs) pack_pairs<<<(unsigned)((npairs + 255) / 256), 256, 0, s>>>(pp);
  if (ncols) pack_cols<<<(unsigned)((ncols + 255) / 256), 256, 0, s>>>(pp);
  GRMP_CUDA(cudaGetLastError());
  if (ntiles > 0) {
    size_t tmp_bytes = 0;
    GRMP_CUDA(cub::DeviceSegmentedSort::SortPairs(nullptr, tmp_bytes, d_mkey.p, d_mkey2.p, d_mval.p, d_mval2.p, (int)nmir_total, ntiles,
                                                  d_mirbase.p, d_mirbase.p + 1, s));
    DevBuf<unsigned char> d_tmp;
    GRMP_TRY(d_tmp.alloc(std::max<size_t>(tmp_bytes, 16)));
    GRMP_CUDA(cub::DeviceSegmentedSort::SortPairs(d_tmp.p, tmp_bytes, d_mkey.p, d_mkey2.p, d_mval.p, d_mval2.p, (int)nmir_total, ntiles,
                                                  d_mirbase.p, d_mirbase.p + 1, s));
    pack_tile_rest<<<ntiles, 128, 0, s>>>(reinterpret_cast<const TileHdr*>(d_hdr.p), d_groups.p, NW, d_nodeids.p, p.g.coords, d_mirbase.p,
                                          d_mkey2.p, d_mval2.p, out->blob.p, out->tile_dir.p);
    GRMP_CUDA(cudaGetLastError());
    GRMP_CUDA(cudaStreamSynchronize(s));      // d_tmp and the sort buffers go out of scope below
  }
  lap("pack kernels + mirror sort");
  // (4) vertex columns: list + diagonal slots
  GRMP_TRY(out->vcols.upload(vcols.data(), vcols.size(), s));
  GRMP_TRY(out->vrec.alloc(2 * vcols.size()));
  if (out->nvcols > 0) {
    find_diag_slots<<<(unsigned)((out->nvcols + 255) / 256), 256, 0, s>>>(out->vcols.p, out->nvcols, pat.colptr.p, pat.rowval.p, d_closed.p, out->vrec.p);
    GRMP_CUDA(cudaGetLastError());
  }
  const int smem_attr = (int)std::max<i64>(max_smem, 1024);
  GRMP_TRY(set_smem_attr<3>(smem_attr)); GRMP_TRY(set_smem_attr<4>(smem_attr)); GRMP_TRY(set_smem_attr<5>(smem_attr));
  GRMP_TRY(set_smem_attr<6>(smem_attr)); GRMP_TRY(set_smem_attr<7>(smem_attr));
  GRMP_CUDA(cudaStreamSynchronize(s));
  return GRMP_OK;
}

template <int NW> int launch_edge(const EdgeParams& ep, FastP2Tet& f, int sm_count, cudaStream_t s, bool pdl) {
  int per_sm = 0;
  GRMP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, p2tet_edge_kernel<NW>, (NW + NSVC) * 32, (size_t)f.smem_bytes));
  if (per_sm < 1) return fail(GRMP_ECUDA, "fast path: edge kernel does not fit on an SM");
  const int grid = std::min(std::min(f.ntiles, per_sm * sm_count), 1024);
  f.grid = grid;
  EdgeParams e2 = ep;
  e2.claims_per_step = (u32)(f.ntiles + grid);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((NW + NSVC) * 32); cfg.dynamicSmemBytes = (size_t)f.smem_bytes; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  GRMP_CUDA(cudaLaunchKernelEx(&cfg, p2tet_edge_kernel<NW>, e2));
  return GRMP_OK;
}

int fast_p2tet_numeric(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, FastP2Tet& f, i64 geom_version, double* nzval) {
  static const int dbg = getenv("GRMP_DEBUG_FLAGS") ? atoi(getenv("GRMP_DEBUG_FLAGS")) : 0;
  // programmatic dependent launch (GRMP_NO_PDL=1 disables it): the edge kernel may overlap the tail of the kernel before it in the
  // stream only if that kernel is the previous step's diagonal kernel -- not when the tile coordinates were just rewritten
  // GRMP_PDL: bit 0 = the diagonal kernel is a programmatic dependent of the edge kernel, bit 1 = the edge kernel of the diagonal kernel before it.
  // Measured (profiles/r2_metric_kernel_experiments.md, 5): bit 1 gains 0.1 % at level 6 and 1.2 % at level 5 (= one rank of eight); bit 0 lets
  // idle diagonal CTAs take the SMs of retired edge CTAs and costs 1.5 % at level 6 -> default 2
  const int pdl_mode = getenv("GRMP_NO_PDL") ? 0 : (getenv("GRMP_PDL") ? atoi(getenv("GRMP_PDL")) : 2);
  const bool use_pdl = (pdl_mode & 1) != 0;
  // ... and only if a diagonal kernel separates consecutive edge kernels (two edge grids must never claim tiles concurrently)
  bool pdl_edge = (pdl_mode & 2) && f.nvcols > 0 && !(dbg & 2);
  if (f.ntiles > 0 && f.geom_version != geom_version) {
    pdl_edge = false;
    repack_tile_coords<<<f.ntiles, 128, 0, ctx->stream>>>(reinterpret_cast<const TileHdr*>(f.tile_hdr.p), f.tile_nodeids.p, p.g.coords, f.blob.p);
    GRMP_CUDA(cudaGetLastError());
    f.geom_version = geom_version;
  }
  if (f.ntiles > 0) {
    EdgeParams ep{f.tile_dir.p, f.blob.p, f.end_slots.p, p.factor, nzval, f.halo_first, f.tile_counter.p, 0u, f.ntiles, f.in_stride, f.slot_elems, f.nbuf, f.prof.p, dbg};
    switch (f.nw) {
      case 3: GRMP_TRY(launch_edge<3>(ep, f, ctx->sm_count, ctx->stream, pdl_edge)); break;
      case 4: GRMP_TRY(launch_edge<4>(ep, f, ctx->sm_count, ctx->stream, pdl_edge)); break;
      case 5: GRMP_TRY(launch_edge<5>(ep, f, ctx->sm_count, ctx->stream, pdl_edge)); break;
      case 6: GRMP_TRY(launch_edge<6>(ep, f, ctx->sm_count, ctx->stream, pdl_edge)); break;
      default: GRMP_TRY(launch_edge<7>(ep, f, ctx->sm_count, ctx->stream, pdl_edge)); break;
    }
    GRMP_CUDA(cudaGetLastError());
  }
  if (f.nvcols > 0 && !(dbg & 2)) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((f.nvcols * 8 + 255) / 256)); cfg.blockDim = dim3(256); cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (use_pdl && f.ntiles > 0) ? 1 : 0;
    GRMP_CUDA(cudaLaunchKernelEx(&cfg, p2tet_vertex_diag_kernel, (const uint4*)f.vrec.p, (i64)f.nvcols, nzval));
  }
  return GRMP_OK;
}

int fast_p2tet_print_prof(grmp_ctx* ctx, const FastP2Tet& f) {
  if (!f.prof.p || f.grid < 1) return GRMP_OK;
  std::vector<unsigned long long> h(f.prof.n);
  GRMP_CUDA(cudaStreamSynchronize(ctx->stream));
  GRMP_CUDA(cudaMemcpy(h.data(), f.prof.p, f.prof.bytes(), cudaMemcpyDeviceToHost));
  double m[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int c = 0; c < f.grid; c++) for (int k = 0; k < 8; k++) m[k] += (double)h[8 * (size_t)c + k] / f.grid;
  fprintf(stderr, "[grmp fast prof] per CTA (mean cycles): service warp wait-done %.0f  mirror write-out %.0f  total %.0f | consumer warp 0 "
                  "wait-full %.0f  total %.0f | tiles %.1f\n", m[0], m[1], m[2], m[3], m[4], m[7]);
  return GRMP_OK;
}

}  // namespace grmp
