// common.cuh -- shared declarations of libgrmp_cuda (handles, device views, error plumbing)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/grmp.h"

namespace grmp {

typedef int32_t i32;
typedef int64_t i64;
typedef uint32_t u32;
typedef uint64_t u64;

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define GRMP_CUDA(call)                                                                          \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      return grmp::fail(e__ == cudaErrorMemoryAllocation ? GRMP_ENOMEM : GRMP_ECUDA,             \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                    \
  } while (0)
#define GRMP_TRY(call)          \
  do {                          \
    int rc__ = (call);          \
    if (rc__ != GRMP_OK) return rc__; \
  } while (0)

// owning device buffer
template <class T> struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  int alloc(size_t count) {
    release();
    if (count == 0) return GRMP_OK;
    GRMP_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));
    n = count;
    return GRMP_OK;
  }
  int upload(const T* host, size_t count, cudaStream_t s) {
    if (n != count) GRMP_TRY(alloc(count));
    if (count) GRMP_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
    return GRMP_OK;
  }
  size_t bytes() const { return n * sizeof(T); }
};

// ---- device views passed to kernels by value -------------------------------------------------
struct GridView {
  int dim;                  // dimension of the items
  int xdim;                 // dimension of the coordinates (> dim: boundary-face items, no affine inverse)
  i64 nnodes, ncells, nfaces;
  const double* coords;     // [nnodes][dim]
  const i32* cellnodes;     // [ncells][dim+1], 1-based
  const double* vol;        // [ncells]
  const i32* regions;       // [ncells] or null
  const i32* cellfaces;     // [ncells][dim+1] 1-based or null
  const i32* signs;         // [ncells][dim+1]
  const i32* orient;        // [ncells][4] (3D) or null
  const double* fnormals;   // [nfaces][dim]
  const double* fvol;       // [nfaces]
};

enum Family { FAM_H1 = 0, FAM_H1BR = 1, FAM_RT0 = 2, FAM_BDM1 = 3 };

// one FEEvaluator (operator applied to a space)
struct EvalView {
  int fam;        // Family of the space
  int op;         // GRMP_OP_*
  int ncomp;      // components of the space
  int nd;         // local dofs
  int nd_all;     // reference functions (BDM1-3D: 16)
  int rd;         // resultdim of the operator
  int rfam;       // reconstruction target family (FAM_RT0 / FAM_BDM1) for RECON ops
  int nd2;        // dofs of the reconstruction space on the cell
  int tab_nd;     // nd_all of the table (reconstruction: of the Hdiv space)
  int tab_nc;     // ncomp of the table
  const double* refvals;    // [nq][tab_nd][tab_nc]
  const double* refderivs;  // [nq][edim][tab_nd*tab_nc] or null
  const i32* celldofs;      // [ncells][nd], 1-based
};

struct RegionFilter {
  int n;            // 0 => all cells (regions == [0])
  i32 r[8];
};

// ---- handles -----------------------------------------------------------------------------------
}  // namespace grmp

struct grmp_ctx {
  int device;
  cudaStream_t stream, copy_stream;   // compute (+ result download) / uploads that may overlap it
  cudaEvent_t ev0, ev1, ev_copy;
  int sm_count;
};

struct grmp_grid {
  grmp_ctx* ctx;
  int dim;                        // dimension of the items (cells; boundary faces for grmp_grid_create_bfaces)
  int xdim;                       // dimension of the coordinates (== dim for cell grids)
  grmp::i64 nnodes, ncells, nfaces;
  grmp::DevBuf<double> coords, vol, fnormals, fvol;
  grmp::DevBuf<grmp::i32> cellnodes, regions, cellfaces, signs, orient;
  bool has_regions = false, has_faces = false;
  grmp::i64 geom_version = 0;     // bumped by grmp_grid_update_geometry (the fast path keeps tile-blocked coordinate copies)
  grmp::i64 topo_version = 0;     // bumped by grmp_grid_update_cells: patterns built before are stale
  grmp::GridView view() const;
};

struct grmp_space {
  grmp_grid* grid;
  int fetype, ncomp, nd;
  grmp::i64 ndofs;
  grmp::DevBuf<grmp::i32> celldofs;
  grmp::i64 topo_version = 0;     // bumped by grmp_space_update_dofs
};

namespace grmp {

struct EvalTables {
  DevBuf<double> refvals, refderivs;
  int nd_all = 0, ncomp = 0;
};

int make_evalview(const grmp_space* sp, int op, const EvalTables& tab, EvalView* out);
int op_resultdim(int op, int ncomp, int edim);

// launchers implemented in generic_kernels.cu (compiled with -fmad=false: reference
// operation order, no contraction -> bit-identical to the un-fused CPU evaluation)
struct BlfLocalParams {
  GridView g;
  EvalView e1, e2;
  int same_eval;       // e2 aliases e1 (assemblypatterns.jl:567-584)
  int action;
  double act_p[2];
  int apt;
  int transposed;      // transposed_assembly
  RegionFilter reg;
  int nq;
  const double* w;     // [nq]
  double factor;
  const double* aq;    // GRMP_ACT_CONVECTION / NEWTON_CONVECTION: a(x_q) of the fixed argument, [ncells][nq][aq_rd]
  int aq_rd;
  EvalView ea;         // NEWTON_CONVECTION: a_operator on the ansatz space (evaluation of the ansatz functions themselves)
  const double* gq;    //   ansatz operator of the current iterate at the quadrature points, [ncells][nq][e1.rd]
  double* rbuf;        //   [nd2][ncells] right-hand side contributions (numeric pass) or null
  unsigned char* active;   // [ncells] region filter result for the right-hand side gather, or null
  i64 nrows_key;       // key = col * nrows_key + row (0-based, output orientation)
  // outputs (exactly one of them non-null)
  u64* keys;           // [ncells*nd1*nd2] : symbolic pass, ~0 for masked-out contributions
  double* lbuf;        // [nd1*nd2][ncells]: numeric pass (value handed to _addnz)
};
int launch_blf_local(const BlfLocalParams& p, cudaStream_t s);

struct LfLocalParams {
  GridView g;
  EvalView e;
  RegionFilter reg;
  int nq;
  const double* w;
  double factor;
  int fsrc;
  const double* fdata;   // device
  double* lbuf;          // [nd][ncells]
  unsigned char* active; // [ncells] 1 if the cell is assembled (region filter)
};
int launch_lf_local(const LfLocalParams& p, cudaStream_t s);

// ItemIntegrator with one argument (itemintegrator.jl:160-300)
struct IiLocalParams {
  GridView g;
  EvalView e;
  RegionFilter reg;
  int nq;
  const double* w;
  int kind;              // GRMP_II_*
  int ardim;             // result length: rd (NONE) or 1
  double factor;         // L2ERROR: || data - factor * discrete ||
  const double* coeffs;  // device, [ndofs] entries of the FEVectorBlock
  const double* data;    // device, [ncells][nq][rd] (L2ERROR)
  double* b;             // device, [ncells][ardim], updated in place (b[j,item] += ...), or null
  double* itemval;       // device, [ardim][ncells]: the item's own sum (0 for filtered cells), for the total
  double* qtable;        // device, [ncells][nq][rd] or null: the operator evaluation itself (fixed arguments of trilinear forms)
};
int launch_ii_local(const IiLocalParams& p, cudaStream_t s);

}  // namespace grmp
