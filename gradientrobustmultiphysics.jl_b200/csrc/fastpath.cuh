// fastpath.cuh -- owner-computes roofline kernel for the metric configuration
// (3D P2 Laplace stiffness, BASELINE.json): see fastpath_p2tet.cu
#pragma once
#include "common.cuh"
#include "symbolic.cuh"

namespace grmp {

struct FastP2Tet {
  int ntiles = 0;
  i64 npairs = 0;
  int smem_bytes = 0;
  int max_tile_cells = 0;
  DevBuf<i32> tile_colbeg;     // [ntiles+1] first (permuted) column of a tile
  DevBuf<i32> tile_cellbeg;    // [ntiles+1] range into tile_cells
  DevBuf<i32> tile_cells;      // distinct cells of every tile
  DevBuf<i64> col_pairbeg;     // [ncols+1] pairs of a column
  DevBuf<uint4> pairs;         // ring-ordered pair records of the edge columns
  DevBuf<uint4> cols;          // 2 per column: fixed-row offsets, closing offsets, mirrored slots
  DevBuf<u32> vcols, vdiag;    // vertex columns and their diagonal slots
  i64 nvcols = 0;
};

bool fast_p2tet_applicable(const BlfLocalParams& p);
int fast_p2tet_build(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, const std::vector<double>& w,
                     const std::vector<double>& derivs, i64 ncols_owned, FastP2Tet* out);
int fast_p2tet_numeric(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, const FastP2Tet& f, double* nzval);

}  // namespace grmp
