// fastpath.cuh -- owner-computes roofline kernel for the metric configuration
// (3D P2 Laplace stiffness, BASELINE.json): see fastpath_p2tet.cu
#pragma once
#include "common.cuh"
#include "symbolic.cuh"

namespace grmp {

struct FastP2Tet {
  int ntiles = 0;
  int tpb = 128;
  i64 npairs = 0;
  int smem_bytes = 0;
  int max_tile_cells = 0;
  DevBuf<int4> tile_hdr;       // 2 per tile: column range, tile-cell range, nzval range
  DevBuf<int4> tile_nodes;     // CellNodes of the distinct cells of every tile
  DevBuf<i64> col_pairbeg;     // [ncols+1] pairs of a column
  DevBuf<uint4> pairs;         // ring-ordered pair records of the edge columns
  DevBuf<uint4> cols;          // 2 per column: fixed-row offsets, closing offsets, mirrored slots
  DevBuf<u32> vcols;           // vertex columns
  DevBuf<uint4> vrec;          // per vertex column: diagonal slot, first spoke slot, #spokes
  DevBuf<uint2> spokes;        // per edge column: scratch slots in the spoke lists of its two end vertices
  DevBuf<double> dscratch;     // per (vertex, spoke): 0.2 * ring sum of S_vv
  i64 nvcols = 0;
};

bool fast_p2tet_applicable(const BlfLocalParams& p);
int fast_p2tet_build(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, const std::vector<double>& w,
                     const std::vector<double>& derivs, i64 ncols_owned, FastP2Tet* out);
int fast_p2tet_numeric(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, const FastP2Tet& f, double* nzval);

}  // namespace grmp
