// fastpath.cuh -- owner-computes roofline kernel for the metric configuration
// (3D P2 Laplace stiffness, BASELINE.json): see fastpath_p2tet.cu
#pragma once
#include "common.cuh"
#include "symbolic.cuh"

namespace grmp {

struct FastP2Tet {
  int ntiles = 0;
  int nw = 7;                     // consumer warps per CTA
  int nbuf = 2;                   // depth of the input ring
  i64 npairs = 0;
  int smem_bytes = 0;
  u32 slot_elems = 0;             // doubles per warp stage slot
  u32 in_stride = 0;              // bytes of one input buffer of the kernel's 2-deep ring (largest blob)
  i64 geom_version = 0;           // grmp_grid::geom_version the blobs' node coordinates were packed from
  DevBuf<int4> tile_hdr;          // 3 per tile (TileHdr), kept for re-packing the coordinates
  DevBuf<u32> tile_nodeids;       // distinct nodes of every tile (1-based), kept for re-packing the coordinates
  DevBuf<uint2> tile_dir;         // per tile: blob offset (16-byte units), blob bytes
  DevBuf<unsigned char> blob;     // per tile: header, group table, column / pair records, node coordinates, sorted mirror list
  DevBuf<u32> end_slots;          // per pair, only when a partition produced multi-chain halo columns
  DevBuf<u32> vcols;              // vertex columns
  DevBuf<uint4> vrec;             // per vertex column: diagonal slot, first slot, #slots
  DevBuf<unsigned long long> tile_counter;   // dynamic tile scheduler of the edge kernel: running claim counter, never reset
  i64 nvcols = 0;
  i64 halo_first = -1;            // first nzval slot of the columns another rank owns (>= nnz: none): never written by the kernels
  DevBuf<unsigned long long> prof; // GRMP_FAST_PROF: cycle counters of the last launch (8 per CTA)
  int grid = 0;                   // CTAs of the last edge-kernel launch
};

bool fast_p2tet_applicable(const BlfLocalParams& p);
// max over cells T and vertices a of sum_{b != a} |S_ab| / S_aa, S = grad(lambda_a).grad(lambda_b): the factor by which the
// row-sum identities of the ring-walk kernel amplify rounding errors (1 on non-obtuse cells)
int fast_p2tet_quality(grmp_ctx* ctx, const BlfLocalParams& p, double* kappa);
int fast_p2tet_build(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, const std::vector<double>& w,
                     const std::vector<double>& derivs, i64 ncols_owned, i64 geom_version, FastP2Tet* out);
// GRMP_FAST_PROF: mean cycle counters of the last launch on stderr
int fast_p2tet_print_prof(grmp_ctx* ctx, const FastP2Tet& f);
int fast_p2tet_numeric(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, FastP2Tet& f, i64 geom_version, double* nzval);

}  // namespace grmp
