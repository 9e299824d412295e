// cellpath.cu -- the cell-parallel scatter alternatives to the owner-computes column kernels (north_star: "scatter-add ...
// under element colouring, which keeps it deterministic; FP64 atomics are a measured alternative").
//
// One thread per cell: geometry and the local matrix live in registers (same evaluator code as the column kernels,
// colpath_ev.cuh, all table indices uniform over the warp), every local entry goes to nzval through the per-cell
// local -> nnz map of the symbolic pass (bilinearform.jl:319-369: rows CellDofs1[di], columns CellDofs2[dj]):
//   mode 0  FP64 atomics (red.global.add.f64) into a zeroed nzval: one launch, summation order not fixed;
//   mode 1  element colouring: cells that share a column dof get different colours, one launch per colour with plain
//           read-modify-write -> deterministic (order = colour order), more launches.
// Both move 16 B of nzval traffic + 4 B of map per local entry (the column kernels: 8 B per stored entry + ~1 B of record per
// local entry), which is why they are the measured alternatives and not the default; numbers in profiles/r2_scatter_variants.md.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "colpath.cuh"
#include "colpath_ev.cuh"

namespace grmp {

namespace {

struct CellParams {
  GridView g;
  const i32* slotmapT;      // [nd1*nd2][ncells]
  const u32* cells;         // cells of this launch (colour) or null = identity
  i64 nlaunch, ncells;
  const double* tabR;       // [q][s][a]
  const double* tabC;       // [a][q][CT_PAD]
  const double* wq;
  double factor;
  double act_p[2];
  double* nzval;
  int nq, mode;
};

template <class RowEv, class ColEv, int ACT>
__global__ void __launch_bounds__(128) cell_kernel(const CellParams p) {
  using L = CacheLayout<RowEv, ColEv>;
  extern __shared__ __align__(16) double sm[];
  const int nq = p.nq;
  const int ntc = ColEv::NAS * nq * CT_PAD, ntr = RowEv::NSF * RowEv::NAS * nq;
  double* const sCt = sm;
  double* const sRt = sm + ntc;
  double* const sW = sRt + ntr;
  for (int i = threadIdx.x; i < ntc; i += blockDim.x) sCt[i] = p.tabC[i];
  for (int i = threadIdx.x; i < ntr; i += blockDim.x) sRt[i] = p.tabR[i];
  for (int i = threadIdx.x; i < nq; i += blockDim.x) sW[i] = p.wq[i];
  __syncthreads();
  const i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (i >= p.nlaunch) return;
  const i64 cell = p.cells ? (i64)p.cells[i] : i;
  double cr[L::STRIDE];
  build_cell_cache<RowEv, ColEv>(p.g, cell, p.factor, cr);
  typename RowEv::Regs RR;
  typename ColEv::Regs RC;
  RowEv::load(cr + L::OFF_R, RR);
  ColEv::load(cr + L::OFF_C, RC);
  const double s = cr[0];
  const i32* __restrict__ sl = p.slotmapT + cell;
#pragma unroll 1
  for (int lc = 0; lc < ColEv::NROW; lc++) {
    typename RowEv::Acc A;
    RowEv::acc_zero(A);
    typename ColEv::Prep PC;
    ColEv::prep(RC, lc, PC);
#pragma unroll 1
    for (int q = 0; q < nq; q++) {
      double Y[ColEv::RD];
      ColEv::col_eval(RC, PC, sCt, nq, q, Y);
      const double ws = sW[q] * s;
#pragma unroll
      for (int k = 0; k < ColEv::RD; k++) Y[k] *= ws;
      apply_action_col<ACT, ColEv::RD>(p.act_p, Y);
      double U[RowEv::NCU][RowEv::NAS];
      RowEv::pullback(RR, Y, U);
      RowEv::acc_rows(A, U, sRt + q * (RowEv::NSF * RowEv::NAS));
    }
    RowEv::emit_rows(RR, A, [&](int r, double v) {
      const i32 slot = sl[(size_t)(r * ColEv::NROW + lc) * p.ncells];
      if (slot >= 0) {
        if (p.mode == 0) atomicAdd(p.nzval + slot, v);
        else p.nzval[slot] += v;
      }
    });
  }
}

__global__ void transpose_slotmap(const i32* slotmap, i64 ncells, int nloc, i32* out) {
  const i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (i >= ncells * nloc) return;
  const i64 cell = i / nloc;
  const int e = (int)(i - cell * nloc);
  out[(size_t)e * ncells + cell] = slotmap[i];
}

typedef int (*CellLaunchFn)(const CellParams&, int smem, cudaStream_t);
struct CellVariant {
  bool (*match)(const ColEvalDesc& row, const ColEvalDesc& col, int act);
  CellLaunchFn launch;
};
template <class RowEv, class ColEv, int ACT> struct CellVariantImpl {
  static bool match(const ColEvalDesc& row, const ColEvalDesc& col, int act) { return act == ACT && ev_matches<RowEv>(row) && ev_matches<ColEv>(col); }
  static int launch(const CellParams& p, int smem, cudaStream_t s) {
    cell_kernel<RowEv, ColEv, ACT><<<(unsigned)((p.nlaunch + 127) / 128), 128, smem, s>>>(p);
    GRMP_CUDA(cudaGetLastError());
    return GRMP_OK;
  }
};
#define GRMP_UNPAREN(...) __VA_ARGS__
#define GRMP_CELLVARIANT(R, C, A, Q) \
  {&CellVariantImpl<GRMP_UNPAREN R, GRMP_UNPAREN C, A>::match, &CellVariantImpl<GRMP_UNPAREN R, GRMP_UNPAREN C, A>::launch},
const CellVariant CELLVARIANTS[] = {GRMP_SQUARE_FORMS(GRMP_CELLVARIANT)};
constexpr int NCELLVARIANTS = sizeof(CELLVARIANTS) / sizeof(CELLVARIANTS[0]);

int find_cell_variant(const ColPath& cp, int action) {
  if (!cp.row_is_arg1) return -1;
  if ((cp.row.kind == 1 && cp.row.nds == 16) || (cp.col.kind == 1 && cp.col.nds == 16)) return -1;   // BDM1 3D: subset selection
  for (int v = 0; v < NCELLVARIANTS; v++)
    if (CELLVARIANTS[v].match(cp.row, cp.col, action)) return v;
  return -1;
}

}  // namespace

int cellpath_build(grmp_ctx* ctx, const BlfLocalParams& p, Pattern& pat, bool coloured, ColPath* cp) {
  cudaStream_t s = ctx->stream;
  if (find_cell_variant(*cp, p.action) < 0) return fail(GRMP_EUNSUPPORTED, "no cell-parallel kernel for this form");
  if (pat.slotmap.n == 0 && pat.nnz > 0) return fail(GRMP_ESTATE, "cell-parallel kernels need the local -> nnz map of the symbolic pass");
  const i64 ncells = p.g.ncells;
  const int nloc = p.e1.nd * p.e2.nd;
  GRMP_TRY(cp->slotmapT.alloc(std::max<size_t>((size_t)ncells * nloc, 1)));
  if (ncells * nloc > 0) {
    transpose_slotmap<<<(unsigned)((ncells * nloc + 255) / 256), 256, 0, s>>>(pat.slotmap.p, ncells, nloc, cp->slotmapT.p);
    GRMP_CUDA(cudaGetLastError());
  }
  GRMP_CUDA(cudaStreamSynchronize(s));
  pat.slotmap.release();
  cp->colour_ptr.clear();
  cp->colour_cells.release();
  if (coloured && ncells > 0) {
    // greedy colouring on the host: two cells conflict iff they share a dof of the COLUMN space (same column is necessary for
    // the same nzval slot)
    const int nd = p.e2.nd;
    std::vector<i32> dofs((size_t)ncells * nd);
    GRMP_CUDA(cudaMemcpyAsync(dofs.data(), p.e2.celldofs, dofs.size() * 4, cudaMemcpyDeviceToHost, s));
    GRMP_CUDA(cudaStreamSynchronize(s));
    i64 ndofs = 0;
    for (i32 d : dofs) ndofs = std::max<i64>(ndofs, d);
    std::vector<i64> ptr(ndofs + 2, 0);
    for (i32 d : dofs) ptr[d + 1]++;
    for (i64 d = 0; d <= ndofs; d++) ptr[d + 1] += ptr[d];
    std::vector<u32> adj(dofs.size());
    {
      std::vector<i64> fill(ptr.begin(), ptr.end() - 1);
      for (i64 c = 0; c < ncells; c++) for (int l = 0; l < nd; l++) adj[fill[dofs[c * nd + l]]++] = (u32)c;
    }
    std::vector<int> colour(ncells, -1);
    std::vector<u64> forb;
    int ncol = 0;
    for (i64 c = 0; c < ncells; c++) {
      forb.assign(16, 0ull);      // up to 1024 colours
      for (int l = 0; l < nd; l++) {
        const i32 d = dofs[c * nd + l];
        for (i64 k = ptr[d]; k < ptr[d + 1]; k++) {
          const int cc = colour[adj[k]];
          if (cc >= 0) forb[cc >> 6] |= 1ull << (cc & 63);
        }
      }
      int pick = -1;
      for (int wd = 0; wd < 16 && pick < 0; wd++)
        if (~forb[wd]) pick = 64 * wd + __builtin_ctzll(~forb[wd]);
      if (pick < 0) return fail(GRMP_EUNSUPPORTED, "colouring needs more than 1024 colours");
      colour[c] = pick;
      ncol = std::max(ncol, pick + 1);
    }
    cp->colour_ptr.assign(ncol + 1, 0);
    for (i64 c = 0; c < ncells; c++) cp->colour_ptr[colour[c] + 1]++;
    for (int k = 0; k < ncol; k++) cp->colour_ptr[k + 1] += cp->colour_ptr[k];
    std::vector<u32> order(ncells);
    {
      std::vector<i64> fill(cp->colour_ptr.begin(), cp->colour_ptr.end() - 1);
      for (i64 c = 0; c < ncells; c++) order[fill[colour[c]]++] = (u32)c;
    }
    GRMP_TRY(cp->colour_cells.upload(order.data(), order.size(), s));
    GRMP_CUDA(cudaStreamSynchronize(s));
  }
  return GRMP_OK;
}

int cellpath_numeric(grmp_ctx* ctx, const BlfLocalParams& p, const Pattern& pat, ColPath& cp, int mode, double* nzval, i64* launches) {
  cudaStream_t s = ctx->stream;
  const int v = find_cell_variant(cp, p.action);
  if (v < 0) return fail(GRMP_EUNSUPPORTED, "no cell-parallel kernel for this form");
  if (launches) *launches = 0;
  if (pat.nnz == 0 || p.g.ncells == 0) return GRMP_OK;
  static DevBuf<double> d_tabR, d_w;     // small per-process scratch: row table + weights of the active form
  static u64 owner = 0;
  if (owner != cp.uid) {
    GRMP_TRY(d_tabR.upload(cp.tabR.data(), cp.tabR.size(), s));
    GRMP_TRY(d_w.upload(cp.wq.data(), cp.wq.size(), s));
    owner = cp.uid;
  }
  CellParams cpar{};
  cpar.g = p.g; cpar.slotmapT = cp.slotmapT.p; cpar.ncells = p.g.ncells; cpar.tabR = d_tabR.p; cpar.tabC = cp.tabC.p; cpar.wq = d_w.p;
  cpar.factor = p.factor; cpar.act_p[0] = p.act_p[0]; cpar.act_p[1] = p.act_p[1]; cpar.nzval = nzval; cpar.nq = cp.nq; cpar.mode = mode;
  const int nsfR = cp.row.nds + cp.row.nbub;
  const int nasR = (int)(cp.tabR.size() / ((size_t)cp.nq * nsfR));
  const int nasC = (int)(cp.tabC.n / ((size_t)cp.nq * CT_PAD));
  const int smem = 8 * (nasC * cp.nq * CT_PAD + nsfR * nasR * cp.nq + cp.nq);
  GRMP_CUDA(cudaMemsetAsync(nzval, 0, (size_t)pat.nnz * 8, s));
  if (mode == 0) {
    cpar.cells = nullptr; cpar.nlaunch = p.g.ncells;
    GRMP_TRY(CELLVARIANTS[v].launch(cpar, smem, s));
    if (launches) *launches = 2;
  } else {
    if (cp.colour_ptr.empty()) return fail(GRMP_ESTATE, "colouring not built");
    const int ncol = (int)cp.colour_ptr.size() - 1;
    for (int c = 0; c < ncol; c++) {
      cpar.cells = cp.colour_cells.p + cp.colour_ptr[c];
      cpar.nlaunch = cp.colour_ptr[c + 1] - cp.colour_ptr[c];
      if (cpar.nlaunch > 0) GRMP_TRY(CELLVARIANTS[v].launch(cpar, smem, s));
    }
    if (launches) *launches = ncol + 1;
  }
  return GRMP_OK;
}

}  // namespace grmp
