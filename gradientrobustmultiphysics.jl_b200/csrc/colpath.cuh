// colpath.cuh -- owner-computes column kernels for every (element, operator) pair on the ported path
// (the kernel family behind GRMP_PATH_COLUMNS) and the cell-parallel scatter alternatives
// (GRMP_PATH_ATOMIC / GRMP_PATH_COLOURED): see colpath.cu
#pragma once
#include "common.cuh"
#include "symbolic.cuh"

namespace grmp {

// One argument of the form as the column kernels see it
struct ColEvalDesc {
  int kind;        // 0: componentwise H1 space (+ BR face bubbles), 1: Hdiv space
  int op;          // GRMP_OP_*
  int ed;          // element dimension
  int nc;          // components (H1) / 1 (Hdiv)
  int nds;         // scalar shape functions per component (H1) / reference functions nd_all (Hdiv)
  int nbub;        // BR face bubbles
  int nd;          // local dofs on the cell
  int fam;         // Family
  bool same(const ColEvalDesc& o) const { return kind == o.kind && ed == o.ed && nc == o.nc && nds == o.nds && nbub == o.nbub && fam == o.fam; }
};

struct ColPath {
  bool built = false;
  u64 uid = 0;                     // identifies whose tables sit in constant memory
  int variant = 0;                 // index into the kernel table (colpath.cu)
  int cf_variant = -1;             // >= 0: closed-form kernel (reference tensors instead of a quadrature loop)
  int nw = 4;                      // warps (= groups of 32 columns) per CTA / tile
  int nv = 1;                      // 16-byte vectors per pair record
  int nq = 0;
  int variant_reduced = -1;        // >= 0: kernel for an equivalent 3-point rule (reconstruction mass forms, see colpath_build)
  i64 ncols_used = 0, ngroups = 0, ntiles = 0, npairs = 0;
  int max_tile_cells = 0, max_grp_nnz = 0;
  int smem_bytes = 0;              // largest class
  // tiles are launched in classes of similar shared-memory need, so that one crowded tile (coarse-level vertices touch many
  // cells) does not cap the occupancy of all the others
  struct TileClass { int smem_bytes; i64 first, count; };
  std::vector<TileClass> classes;
  DevBuf<u32> class_tiles;         // tile ids ordered by class
  ColEvalDesc row{}, col{};
  bool row_is_arg1 = true;         // rows of the output = first argument (no transposed_assembly)
  DevBuf<uint4> recs;              // pair records, round-major inside every group of 32 columns
  DevBuf<u32> colperm;             // position in the locality order -> column
  DevBuf<unsigned short> pos_np;   // per position: pairs (cells) of the column
  DevBuf<unsigned char> pos_len;   //   stored entries of the column
  DevBuf<i64> pos_start;           //   first nzval slot of the column (0-based)
  DevBuf<i64> pos_recbeg;          // [ncols_used+1] first record of every position (groups start at multiples of 32)
  DevBuf<double> geo;              // [ncells][stride] per-cell geometry records, rewritten by every numeric call
  DevBuf<double> tabC;             // column-function table [a][q][16]
  std::vector<double> tabR;        // row table [s][a][q] -> constant memory at launch
  std::vector<double> wq;
  // cell-parallel alternatives (J2)
  DevBuf<i32> slotmapT;            // [nd_row*nd_col][ncells] local -> nnz (-1: not stored)
  DevBuf<u32> colour_cells;        // cells ordered by colour
  std::vector<i64> colour_ptr;     // [ncolours+1]
};

// LinearForm: one thread per dof gathers the contributions of the cells around it (register accumulation, cells ascending)
struct LfPath {
  bool built = false;
  int variant = -1, nw = 4, nq = 0;
  i64 ndofs = 0, ngroups = 0, ntiles = 0, npairs = 0;
  ColEvalDesc ev{};
  DevBuf<u32> recs;                // per pair: tile-local cell | local function << 16 (255: cell filtered out by the regions)
  DevBuf<u32> dofperm;             // position -> dof
  DevBuf<unsigned short> pos_np;
  DevBuf<i64> pos_recbeg;
  DevBuf<u32> tile_cellptr, tile_cells, class_tiles;
  std::vector<ColPath::TileClass> classes;
  DevBuf<double> tabC, wq;
};
bool lfpath_applicable(const EvalView& e, int edim, int nq, LfPath* lp);
int lfpath_build(grmp_ctx* ctx, const GridView& g, const EvalView& e, const RegionFilter& reg, const std::vector<double>& w,
                 const std::vector<double>& vals, const std::vector<double>& derivs, i64 ndofs, LfPath* lp);
// b[dof] += sum over the cells of the dof, cells ascending (linearform.jl:181-220); fdata: device pointer (GRMP_F_CONST: rd values,
// GRMP_F_QP_TABLE: [ncells][nq][rd])
int lfpath_numeric(grmp_ctx* ctx, const GridView& g, LfPath& lp, double factor, int fsrc, const double* fdata, double* b);

// does a column kernel exist for this form?  (fills the descriptors)
bool colpath_applicable(const BlfLocalParams& p, int nq, ColPath* cp);
// one-time build of the records (device), tables (host copies of the caller's tables) ...
int colpath_build(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, const std::vector<double>& w,
                  const std::vector<double>& vals1, const std::vector<double>& derivs1, const std::vector<double>& vals2,
                  const std::vector<double>& derivs2, i64 ncols_owned, bool quadrature_tables, ColPath* cp);
// ... and of the per-cell local -> nnz map (+ greedy element colouring when `coloured`) for the cell-parallel kernels
int cellpath_build(grmp_ctx* ctx, const BlfLocalParams& p, Pattern& pat, bool coloured, ColPath* cp);
int colpath_numeric(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, ColPath& cp, double* nzval);
// mode 0: FP64 atomics (nzval zeroed first), 1: one launch per colour, plain read-modify-write
int cellpath_numeric(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, ColPath& cp, int mode, double* nzval, i64* launches);

}  // namespace grmp
