// api.cu -- the C ABI of libgrmp_cuda (include/grmp.h): handles, uploads, dispatch.
#include <cstring>
#include <mutex>

#include "colpath.cuh"
#include "common.cuh"
#include "fastpath.cuh"
#include "sparseops.cuh"
#include "symbolic.cuh"

namespace grmp {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

int op_resultdim(int op, int ncomp, int edim) {   // Length4Operator, src/functionoperators.jl:260-278
  switch (op) {
    case GRMP_OP_ID: case GRMP_OP_RECON_ID_RT0: case GRMP_OP_RECON_ID_BDM1: return ncomp;
    case GRMP_OP_GRAD: return edim * ncomp;
    case GRMP_OP_SYMGRAD: return ((edim == 2) ? 3 : 6) * ((ncomp + edim - 1) / edim);
    case GRMP_OP_DIV: return (ncomp + edim - 1) / edim;
    case GRMP_OP_NORMALFLUX: return 1;
  }
  return -1;
}

static int fe_family(int fetype) {
  switch (fetype) {
    case GRMP_FE_H1P1: case GRMP_FE_H1P2: case GRMP_FE_L2P0: return FAM_H1;
    case GRMP_FE_H1BR: return FAM_H1BR;
    case GRMP_FE_HDIVRT0: return FAM_RT0;
    case GRMP_FE_HDIVBDM1: return FAM_BDM1;
  }
  return -1;
}

static int fe_local_dofs(int fetype, int ncomp, int edim, int* nd, int* nd_all, int* ncomp_eff, bool on_faces = false) {
  if (on_faces && (fetype == GRMP_FE_HDIVRT0 || fetype == GRMP_FE_HDIVBDM1)) {   // normal-flux face bases: scalar valued, "i1" / "i2" / "i3"
    *ncomp_eff = 1;
    *nd = *nd_all = (fetype == GRMP_FE_HDIVRT0) ? 1 : edim + 1;
    return GRMP_OK;
  }
  const int nn = edim + 1, nf = edim + 1, ne = (edim == 1) ? 1 : (edim == 2) ? 3 : 6;   // Edge1D: the interior dof takes the edge slot ("N1I1")
  *ncomp_eff = ncomp;
  switch (fetype) {
    case GRMP_FE_H1P1: *nd = *nd_all = nn * ncomp; return GRMP_OK;
    case GRMP_FE_H1P2: *nd = *nd_all = (nn + ne) * ncomp; return GRMP_OK;
    case GRMP_FE_H1BR: *ncomp_eff = edim; *nd = *nd_all = nf + nn * edim; return GRMP_OK;
    case GRMP_FE_HDIVRT0: *ncomp_eff = edim; *nd = *nd_all = nf; return GRMP_OK;
    case GRMP_FE_HDIVBDM1: *ncomp_eff = edim; *nd = edim * nf; *nd_all = (edim == 2) ? 2 * nf : 4 * nf; return GRMP_OK;
    case GRMP_FE_L2P0: *nd = *nd_all = ncomp; return GRMP_OK;
  }
  return fail(GRMP_EUNSUPPORTED, "unknown FEType code");
}

int make_evalview(const grmp_space* sp, int op, const EvalTables& tab, EvalView* out) {
  const int edim = sp->grid->dim;
  const bool on_faces = sp->grid->xdim != edim;
  EvalView e{};
  int nd, nd_all, nc;
  GRMP_TRY(fe_local_dofs(sp->fetype, sp->ncomp, edim, &nd, &nd_all, &nc, on_faces));
  e.fam = fe_family(sp->fetype);
  e.op = op; e.ncomp = nc; e.nd = nd; e.nd_all = nd_all;
  e.rd = op_resultdim(op, nc, edim);
  if (e.rd < 0) return fail(GRMP_EUNSUPPORTED, "unknown operator code");
  if (e.rd > 9) return fail(GRMP_EUNSUPPORTED, "operator result dimension > 9");
  const bool hdiv_fam = (e.fam == FAM_RT0 || e.fam == FAM_BDM1);
  if (on_faces && !((e.fam == FAM_H1 && op == GRMP_OP_ID) || (hdiv_fam && op == GRMP_OP_NORMALFLUX)))
    return fail(GRMP_EUNSUPPORTED, "boundary-face grids: Identity of H1P1 / H1P2 and NormalFlux of HDIVRT0 / HDIVBDM1 only");
  if (!on_faces && op == GRMP_OP_NORMALFLUX) return fail(GRMP_EUNSUPPORTED, "NormalFlux lives on boundary-face grids (grmp_grid_create_bfaces)");
  const bool hdiv = (e.fam == FAM_RT0 || e.fam == FAM_BDM1);
  if (hdiv && !on_faces && !(op == GRMP_OP_ID || op == GRMP_OP_DIV)) return fail(GRMP_EUNSUPPORTED, "Hdiv elements: Identity / Divergence only");
  if (sp->fetype == GRMP_FE_L2P0 && op != GRMP_OP_ID) return fail(GRMP_EUNSUPPORTED, "L2P0: Identity only");
  if (op == GRMP_OP_SYMGRAD && nc != edim) return fail(GRMP_EINVAL, "SymmetricGradient requires ncomponents == dim");
  const bool recon = (op == GRMP_OP_RECON_ID_RT0 || op == GRMP_OP_RECON_ID_BDM1);
  if (recon && sp->fetype != GRMP_FE_H1BR) return fail(GRMP_EUNSUPPORTED, "ReconstructionIdentity is ported for H1BR only");
  if ((hdiv || sp->fetype == GRMP_FE_H1BR) && !on_faces && !sp->grid->has_faces)
    return fail(GRMP_ESTATE, "grid face data missing: call grmp_grid_set_faces first");
  if (e.fam == FAM_BDM1 && !on_faces && edim == 3 && sp->grid->orient.n == 0) return fail(GRMP_ESTATE, "CellFaceOrientations missing (BDM1 3D)");
  if (recon && op == GRMP_OP_RECON_ID_BDM1 && edim == 3 && sp->grid->orient.n == 0)
    return fail(GRMP_ESTATE, "CellFaceOrientations missing (BR->BDM1 3D)");
  e.tab_nd = nd_all; e.tab_nc = nc;
  if (recon) {
    e.rfam = (op == GRMP_OP_RECON_ID_RT0) ? FAM_RT0 : FAM_BDM1;
    int nd2, nd2_all, nc2;
    GRMP_TRY(fe_local_dofs(op == GRMP_OP_RECON_ID_RT0 ? GRMP_FE_HDIVRT0 : GRMP_FE_HDIVBDM1, edim, edim, &nd2, &nd2_all, &nc2));
    e.nd2 = nd2; e.tab_nd = nd2_all; e.tab_nc = nc2;
  }
  if (tab.nd_all != e.tab_nd || tab.ncomp != e.tab_nc) return fail(GRMP_EINVAL, "evaluator table shape does not match FEType/operator");
  const bool need_deriv = (op == GRMP_OP_GRAD || op == GRMP_OP_SYMGRAD || op == GRMP_OP_DIV);
  if (need_deriv && tab.refderivs.n == 0) return fail(GRMP_EINVAL, "operator needs refderivs");
  if (!need_deriv && tab.refvals.n == 0) return fail(GRMP_EINVAL, "operator needs refvals");
  e.refvals = tab.refvals.p; e.refderivs = tab.refderivs.p;
  e.celldofs = sp->celldofs.p;
  *out = e;
  return GRMP_OK;
}

static int upload_tables(const grmp_evaltab* t, int nq, int edim, cudaStream_t s, EvalTables* out) {
  if (!t) return fail(GRMP_EINVAL, "evaluator table missing");
  out->nd_all = t->nd_all; out->ncomp = t->ncomp;
  if (t->refvals) GRMP_TRY(out->refvals.upload(t->refvals, (size_t)nq * t->nd_all * t->ncomp, s));
  if (t->refderivs) GRMP_TRY(out->refderivs.upload(t->refderivs, (size_t)nq * edim * t->nd_all * t->ncomp, s));
  return GRMP_OK;
}

static int make_regions(const int32_t* regions, int nregions, RegionFilter* r) {
  r->n = 0;
  if (nregions <= 0 || !regions || (nregions == 1 && regions[0] == 0)) return GRMP_OK;   // regions == [0]
  if (nregions > 8) return fail(GRMP_EUNSUPPORTED, "more than 8 regions");
  r->n = nregions;
  for (int k = 0; k < nregions; k++) r->r[k] = regions[k];
  return GRMP_OK;
}

}  // namespace grmp

using namespace grmp;

GridView grmp_grid::view() const {
  GridView v{};
  v.dim = dim; v.xdim = xdim; v.nnodes = nnodes; v.ncells = ncells; v.nfaces = nfaces;
  v.coords = coords.p; v.cellnodes = cellnodes.p; v.vol = vol.p; v.regions = has_regions ? regions.p : nullptr;
  v.cellfaces = cellfaces.p; v.signs = signs.p; v.orient = orient.p; v.fnormals = fnormals.p; v.fvol = fvol.p;
  return v;
}

struct grmp_blf {
  grmp_space *s1, *s2;
  int op1, op2, action, apt, transposed, nq, path_req = GRMP_PATH_AUTO, path = 0;
  double act_p[2] = {0, 0};
  RegionFilter reg;
  bool same_eval;
  DevBuf<double> w;
  EvalTables t1, t2;
  std::vector<double> w_host;
  std::vector<double> t1_derivs_host, t1_vals_host, t2_derivs_host, t2_vals_host;   // host copies of the caller's evaluator tables
  Pattern pat;
  ColPath colp;
  CsrView csr;
  i64 topo_grid = 0, topo_s1 = 0, topo_s2 = 0;     // versions of CellNodes / CellDofs the pattern was built from
  i64 halo_first_slot = -1;                        // first nzval slot of the columns another rank owns (-1: none)
  DevBuf<i32> cmp_nodes, cmp_dofs1, cmp_dofs2;      // scratch of grmp_blf_assemble_host's topology check
  DevBuf<int> cmp_flag;
  DevBuf<double> vec_in, vec_out, vec_rhs, scalar;    // vectors of the matrix-vector entry points
  DevBuf<i64> fixed;
  double last_matmul_ms = 0.0;
  bool have_pattern = false, have_values = false;
  DevBuf<double> lbuf, nzval;
  grmp_space* sa = nullptr;       // fixed (coefficient) argument of a trilinear form
  int opa = 0, aq_rd = 0;
  EvalTables ta;
  DevBuf<double> acoeffs, aq, aq_scratch, gq, rbuf, rhs;
  DevBuf<unsigned char> nl_active;
  DofGather nl_dg;
  bool nl_dg_built = false;
  FastP2Tet fast;
  i64 ncols_owned = -1;
  double p2tet_kappa = 0.0;       // cancellation indicator of the grid (AUTO guard of the ring-walk kernel)
  grmp_stats st{};
  i64 out_rows() const { return (apt != GRMP_APT_SYMMETRIC && transposed) ? s2->ndofs : s1->ndofs; }
  i64 out_cols() const { return (apt != GRMP_APT_SYMMETRIC && transposed) ? s1->ndofs : s2->ndofs; }
};

struct grmp_lf {
  grmp_space* sp;
  int op, nq, path_req = GRMP_PATH_AUTO, path = GRMP_PATH_GENERIC;
  bool fast_tried = false;
  LfPath lfp;
  std::vector<double> w_host, vals_host, derivs_host;
  i64 topo_grid = 0, topo_sp = 0;
  RegionFilter reg;
  DevBuf<double> w, lbuf, b, fdata, acoeffs, ascratch;
  DevBuf<unsigned char> active;
  EvalTables tab, ta;
  grmp_space* sa = nullptr;      // coefficient argument of grmp_lf_assemble_feb
  int opa = 0;
  DofGather dg;
  grmp_stats st{};
};

struct grmp_ii {
  grmp_space* sp;
  int op, kind, nq;
  i64 topo_grid = 0, topo_sp = 0;
  RegionFilter reg;
  DevBuf<double> w, coeffs, data, b, itemval, total;
  DevBuf<unsigned char> tmp;
  EvalTables tab;
};

static int blf_numeric_launch(grmp_blf* b, BlfLocalParams& p, cudaStream_t s);

static int fill_blf_params(grmp_blf* b, double factor, BlfLocalParams* p) {
  p->g = b->s1->grid->view();
  GRMP_TRY(make_evalview(b->s1, b->op1, b->t1, &p->e1));
  if (b->same_eval) p->e2 = p->e1; else GRMP_TRY(make_evalview(b->s2, b->op2, b->t2, &p->e2));
  p->same_eval = b->same_eval ? 1 : 0;
  p->action = b->action; p->act_p[0] = b->act_p[0]; p->act_p[1] = b->act_p[1];
  p->apt = b->apt; p->transposed = b->transposed; p->reg = b->reg; p->nq = b->nq; p->w = b->w.p; p->factor = factor;
  p->nrows_key = b->out_rows();
  p->aq = b->aq.p; p->aq_rd = b->aq_rd;
  p->gq = b->gq.p; p->rbuf = nullptr; p->active = nullptr;
  if (b->action == GRMP_ACT_CONVECTION && (!b->sa || !b->aq.p)) return fail(GRMP_ESTATE, "GRMP_ACT_CONVECTION: call grmp_blf_set_fixed_argument first");
  if (b->action == GRMP_ACT_NEWTON_CONVECTION) {
    if (!b->sa || !b->aq.p || !b->gq.p) return fail(GRMP_ESTATE, "GRMP_ACT_NEWTON_CONVECTION: call grmp_blf_set_newton_argument first");
    GRMP_TRY(make_evalview(b->s1, b->opa, b->ta, &p->ea));
  }
  p->keys = nullptr; p->lbuf = nullptr;
  return GRMP_OK;
}

static int blf_check_fresh(grmp_blf* b) {
  if (!b->have_pattern) return fail(GRMP_ESTATE, "grmp_blf_symbolic has not been called");
  if (b->topo_grid != b->s1->grid->topo_version || b->topo_s1 != b->s1->topo_version || b->topo_s2 != b->s2->topo_version)
    return fail(GRMP_ESTATE, "CellNodes / CellDofs changed after grmp_blf_symbolic: the pattern is stale, run the symbolic pass again");
  return GRMP_OK;
}

__global__ void compare_i32(const i32* a, const i32* b, i64 n, int* differs) {
  const i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (i < n && a[i] != b[i]) *differs = 1;
}

static int blf_numeric_launch(grmp_blf* b, BlfLocalParams& p, cudaStream_t s) {
  grmp_ctx* ctx = b->s1->grid->ctx;
  if (b->path == GRMP_PATH_FAST) {
    GRMP_TRY(fast_p2tet_numeric(ctx, p, b->pat, b->fast, b->s1->grid->geom_version, b->nzval.p));
    b->st.kernel_launches = 2;
    // halo columns (owned by another rank) are walked for their mirrors only; the kernels never store into their slots, which
    // were zeroed once by the symbolic pass
  } else if (b->path == GRMP_PATH_COLUMNS) {
    GRMP_TRY(colpath_numeric(ctx, p, b->pat, b->colp, b->nzval.p));
    b->st.kernel_launches = 1 + (i64)b->colp.classes.size();    // geometry records + one launch per tile class
  } else if (b->path == GRMP_PATH_ATOMIC || b->path == GRMP_PATH_COLOURED) {
    i64 nl = 0;
    GRMP_TRY(cellpath_numeric(ctx, p, b->pat, b->colp, b->path == GRMP_PATH_ATOMIC ? 0 : 1, b->nzval.p, &nl));
    b->st.kernel_launches = nl;
  } else {
    const size_t nl = (size_t)p.e1.nd * p.e2.nd * p.g.ncells;
    if (b->lbuf.n != nl) GRMP_TRY(b->lbuf.alloc(nl));
    if (b->action == GRMP_ACT_NEWTON_CONVECTION) {
      const size_t nr = (size_t)std::max<i64>((i64)p.e2.nd * p.g.ncells, 1);
      if (b->rbuf.n != nr) GRMP_TRY(b->rbuf.alloc(nr));
      if (b->nl_active.n != (size_t)std::max<i64>(p.g.ncells, 1)) GRMP_TRY(b->nl_active.alloc((size_t)std::max<i64>(p.g.ncells, 1)));
      p.rbuf = b->rbuf.p; p.active = b->nl_active.p;
    }
    p.lbuf = b->lbuf.p;
    GRMP_TRY(launch_blf_local(p, s));
    GRMP_TRY(launch_gather(s, b->pat, b->lbuf.p, b->nzval.p));
    b->st.kernel_launches = 2;
  }
  return GRMP_OK;
}

extern "C" {

const char* grmp_last_error(void) { return g_err.c_str(); }

int grmp_init(int device, grmp_ctx** out) {
  if (!out) return fail(GRMP_EINVAL, "out == NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(GRMP_ECUDA, std::string("no CUDA device available (the assembly path has no CPU fallback): ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(GRMP_EINVAL, "device index out of range");
  GRMP_CUDA(cudaSetDevice(device));
  grmp_ctx* c = new grmp_ctx();
  c->device = device;
  GRMP_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  GRMP_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  GRMP_CUDA(cudaEventCreate(&c->ev0));
  GRMP_CUDA(cudaEventCreate(&c->ev1));
  GRMP_CUDA(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
  cudaDeviceProp prop;
  GRMP_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  *out = c;
  return GRMP_OK;
}

int grmp_finalize(grmp_ctx* ctx) {
  if (!ctx) return GRMP_OK;
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->copy_stream);
  cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1); cudaEventDestroy(ctx->ev_copy);
  cudaStreamDestroy(ctx->stream); cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
  return GRMP_OK;
}

int grmp_device_synchronize(grmp_ctx* ctx) {
  if (!ctx) return fail(GRMP_EINVAL, "ctx == NULL");
  GRMP_CUDA(cudaStreamSynchronize(ctx->stream));
  return GRMP_OK;
}

int grmp_grid_create(grmp_ctx* ctx, int dim, int64_t nnodes, const double* coords, int64_t ncells, const int32_t* cellnodes,
                     const double* cellvolumes, const int32_t* cellregions, grmp_grid** out) {
  if (!ctx || !out || !coords || !cellnodes || !cellvolumes) return fail(GRMP_EINVAL, "grmp_grid_create: NULL argument");
  if (dim != 2 && dim != 3) return fail(GRMP_EUNSUPPORTED, "only Triangle2D / Tetrahedron3D grids are on the ported path");
  if (nnodes < 0 || ncells < 0) return fail(GRMP_EINVAL, "negative size");
  GRMP_CUDA(cudaSetDevice(ctx->device));
  grmp_grid* g = new grmp_grid();
  g->ctx = ctx; g->dim = dim; g->xdim = dim; g->nnodes = nnodes; g->ncells = ncells; g->nfaces = 0;
  int rc = g->coords.upload(coords, (size_t)nnodes * dim, ctx->stream);
  if (!rc) rc = g->cellnodes.upload(cellnodes, (size_t)ncells * (dim + 1), ctx->stream);
  if (!rc) rc = g->vol.upload(cellvolumes, (size_t)ncells, ctx->stream);
  if (!rc && cellregions) { rc = g->regions.upload(cellregions, (size_t)ncells, ctx->stream); g->has_regions = true; }
  if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(GRMP_ECUDA, "upload failed");
  if (rc) { delete g; return rc; }
  *out = g;
  return GRMP_OK;
}

// ON_BFACES assembly (src/assemblypatterns.jl:400-440 with AT = ON_BFACES; boundarydata.jl:297-347): the boundary faces as items
int grmp_grid_create_bfaces(grmp_ctx* ctx, int xdim, int64_t nnodes, const double* coords, int64_t nbfaces, const int32_t* bfacenodes,
                            const double* bfacevolumes, const int32_t* bfaceregions, grmp_grid** out) {
  if (!ctx || !out || !coords || !bfacenodes || !bfacevolumes) return fail(GRMP_EINVAL, "grmp_grid_create_bfaces: NULL argument");
  if (xdim != 2 && xdim != 3) return fail(GRMP_EUNSUPPORTED, "boundary faces of Triangle2D / Tetrahedron3D grids only");
  if (nnodes < 0 || nbfaces < 0) return fail(GRMP_EINVAL, "negative size");
  GRMP_CUDA(cudaSetDevice(ctx->device));
  grmp_grid* g = new grmp_grid();
  const int edim = xdim - 1;
  g->ctx = ctx; g->dim = edim; g->xdim = xdim; g->nnodes = nnodes; g->ncells = nbfaces; g->nfaces = 0;
  int rc = g->coords.upload(coords, (size_t)nnodes * xdim, ctx->stream);
  if (!rc) rc = g->cellnodes.upload(bfacenodes, (size_t)nbfaces * (edim + 1), ctx->stream);
  if (!rc) rc = g->vol.upload(bfacevolumes, (size_t)nbfaces, ctx->stream);
  if (!rc && bfaceregions) { rc = g->regions.upload(bfaceregions, (size_t)nbfaces, ctx->stream); g->has_regions = true; }
  if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(GRMP_ECUDA, "upload failed");
  if (rc) { delete g; return rc; }
  *out = g;
  return GRMP_OK;
}

int grmp_grid_set_faces(grmp_grid* g, int64_t nfaces, const int32_t* cellfaces, const int32_t* cellfacesigns,
                        const int32_t* cellfaceorient, const double* facenormals, const double* facevolumes) {
  if (!g || !cellfaces || !cellfacesigns || !facenormals || !facevolumes) return fail(GRMP_EINVAL, "grmp_grid_set_faces: NULL argument");
  if (g->xdim != g->dim) return fail(GRMP_EUNSUPPORTED, "boundary-face grids carry no face data");
  cudaStream_t s = g->ctx->stream;
  const size_t nf = g->dim + 1;
  GRMP_TRY(g->cellfaces.upload(cellfaces, (size_t)g->ncells * nf, s));
  GRMP_TRY(g->signs.upload(cellfacesigns, (size_t)g->ncells * nf, s));
  if (cellfaceorient) GRMP_TRY(g->orient.upload(cellfaceorient, (size_t)g->ncells * nf, s));
  GRMP_TRY(g->fnormals.upload(facenormals, (size_t)nfaces * g->dim, s));
  GRMP_TRY(g->fvol.upload(facevolumes, (size_t)nfaces, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  g->nfaces = nfaces; g->has_faces = true;
  return GRMP_OK;
}

int grmp_grid_update_geometry(grmp_grid* g, const double* coords, const double* cellvolumes) {
  if (!g || !coords || !cellvolumes) return fail(GRMP_EINVAL, "grmp_grid_update_geometry: NULL argument");
  cudaStream_t s = g->ctx->stream;
  GRMP_TRY(g->coords.upload(coords, (size_t)g->nnodes * g->xdim, s));
  GRMP_TRY(g->vol.upload(cellvolumes, (size_t)g->ncells, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  g->geom_version++;
  return GRMP_OK;
}

int grmp_grid_update_cells(grmp_grid* g, const int32_t* cellnodes) {
  if (!g || !cellnodes) return fail(GRMP_EINVAL, "grmp_grid_update_cells: NULL argument");
  cudaStream_t s = g->ctx->stream;
  GRMP_TRY(g->cellnodes.upload(cellnodes, (size_t)g->ncells * (g->dim + 1), s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  g->topo_version++;
  g->geom_version++;
  return GRMP_OK;
}

int grmp_grid_destroy(grmp_grid* g) { delete g; return GRMP_OK; }

int grmp_space_create(grmp_grid* grid, int fetype, int ncomp, int64_t ndofs, int nd_cell, const int32_t* celldofs, grmp_space** out) {
  if (!grid || !celldofs || !out) return fail(GRMP_EINVAL, "grmp_space_create: NULL argument");
  int nd, nd_all, nc;
  GRMP_TRY(fe_local_dofs(fetype, ncomp, grid->dim, &nd, &nd_all, &nc, grid->xdim != grid->dim));
  if (nd != nd_cell) return fail(GRMP_EINVAL, "nd_cell does not match the FEType on this geometry");
  grmp_space* s = new grmp_space();
  s->grid = grid; s->fetype = fetype; s->ncomp = ncomp; s->nd = nd; s->ndofs = ndofs;
  int rc = s->celldofs.upload(celldofs, (size_t)grid->ncells * nd, grid->ctx->stream);
  if (!rc && cudaStreamSynchronize(grid->ctx->stream) != cudaSuccess) rc = fail(GRMP_ECUDA, "upload failed");
  if (rc) { delete s; return rc; }
  *out = s;
  return GRMP_OK;
}
int grmp_space_update_dofs(grmp_space* sp, const int32_t* celldofs) {
  if (!sp || !celldofs) return fail(GRMP_EINVAL, "grmp_space_update_dofs: NULL argument");
  cudaStream_t s = sp->grid->ctx->stream;
  GRMP_TRY(sp->celldofs.upload(celldofs, (size_t)sp->grid->ncells * sp->nd, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  sp->topo_version++;
  return GRMP_OK;
}
int grmp_space_destroy(grmp_space* s) { delete s; return GRMP_OK; }

int grmp_blf_create(grmp_space* s1, grmp_space* s2, int op1, int op2, int action, const double* act_params, int apt,
                    int transposed_assembly, const int32_t* regions, int nregions, int nq, const double* qweights,
                    const grmp_evaltab* tab1, const grmp_evaltab* tab2, grmp_blf** out) {
  if (!s1 || !s2 || !qweights || !tab1 || !out || nq <= 0) return fail(GRMP_EINVAL, "grmp_blf_create: bad argument");
  if (s1->grid != s2->grid) return fail(GRMP_EINVAL, "spaces live on different grids");
  if (apt < 0 || apt > 2) return fail(GRMP_EINVAL, "unknown assembly pattern type");
  if (action < GRMP_ACT_NONE || action > GRMP_ACT_NEWTON_CONVECTION) return fail(GRMP_EINVAL, "unknown action");
  if (action == GRMP_ACT_NEWTON_CONVECTION && (apt != GRMP_APT_BILINEARFORM || s1 != s2)) return fail(GRMP_EINVAL, "the Newton convection form lives on one space and is a general form");
  if ((action == GRMP_ACT_HOOKE2D || action == GRMP_ACT_HOOKE3D) && !act_params) return fail(GRMP_EINVAL, "action parameters missing");
  if (action == GRMP_ACT_CONVECTION && apt != GRMP_APT_BILINEARFORM) return fail(GRMP_EINVAL, "the convection form is a general BilinearForm");
  grmp_blf* b = new grmp_blf();
  cudaStream_t st = s1->grid->ctx->stream;
  b->s1 = s1; b->s2 = s2; b->op1 = op1; b->op2 = op2; b->action = action; b->apt = apt; b->transposed = transposed_assembly ? 1 : 0;
  b->nq = nq;
  if (act_params) { b->act_p[0] = act_params[0]; b->act_p[1] = act_params[1]; }
  b->same_eval = (s1 == s2 && op1 == op2);
  int rc = make_regions(regions, nregions, &b->reg);
  if (!rc) rc = b->w.upload(qweights, nq, st);
  b->w_host.assign(qweights, qweights + nq);
  if (!rc) rc = upload_tables(tab1, nq, s1->grid->dim, st, &b->t1);
  if (!rc && tab1->refderivs) b->t1_derivs_host.assign(tab1->refderivs, tab1->refderivs + (size_t)nq * s1->grid->dim * tab1->nd_all * tab1->ncomp);
  if (!rc && tab1->refvals) b->t1_vals_host.assign(tab1->refvals, tab1->refvals + (size_t)nq * tab1->nd_all * tab1->ncomp);
  if (!rc && !b->same_eval) {
    const grmp_evaltab* t2 = tab2 ? tab2 : tab1;
    rc = upload_tables(t2, nq, s1->grid->dim, st, &b->t2);
    if (!rc && t2->refderivs) b->t2_derivs_host.assign(t2->refderivs, t2->refderivs + (size_t)nq * s1->grid->dim * t2->nd_all * t2->ncomp);
    if (!rc && t2->refvals) b->t2_vals_host.assign(t2->refvals, t2->refvals + (size_t)nq * t2->nd_all * t2->ncomp);
  }
  BlfLocalParams p;
  if (!rc && (action == GRMP_ACT_CONVECTION || action == GRMP_ACT_NEWTON_CONVECTION)) {   // validated when the fixed argument arrives
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = fail(GRMP_ECUDA, "upload failed");
    if (rc) { delete b; return rc; }
    *out = b;
    return GRMP_OK;
  }
  if (!rc) rc = fill_blf_params(b, 1.0, &p);   // validates the combination
  if (!rc) {
    const int ar = (action == GRMP_ACT_NONE) ? p.e1.rd : (action == GRMP_ACT_HOOKE2D ? 3 : 6);
    if (action != GRMP_ACT_NONE && p.e1.rd != ar) rc = fail(GRMP_EINVAL, "action input size does not match the operator");
    else if (ar != p.e2.rd) rc = fail(GRMP_EINVAL, "operator result dimensions do not match");
    else if (apt == GRMP_APT_SYMMETRIC && p.e1.nd != p.e2.nd) rc = fail(GRMP_EINVAL, "symmetric form needs equal local dof counts");
  }
  if (!rc && cudaStreamSynchronize(st) != cudaSuccess) rc = fail(GRMP_ECUDA, "upload failed");
  if (rc) { delete b; return rc; }
  *out = b;
  return GRMP_OK;
}

int grmp_blf_destroy(grmp_blf* b) { delete b; return GRMP_OK; }

int grmp_blf_set_newton_argument(grmp_blf* b, int op_a, const grmp_evaltab* tab_a, const double* coeffs_host, int keep_pattern) {
  if (!b || !tab_a || !coeffs_host) return fail(GRMP_EINVAL, "grmp_blf_set_newton_argument: NULL argument");
  if (b->action != GRMP_ACT_NEWTON_CONVECTION) return fail(GRMP_EINVAL, "the form was not created with GRMP_ACT_NEWTON_CONVECTION");
  grmp_space* su = b->s1;
  grmp_ctx* ctx = su->grid->ctx;
  cudaStream_t s = ctx->stream;
  GRMP_CUDA(cudaSetDevice(ctx->device));
  if (b->sa != su || b->opa != op_a || (!b->ta.refvals.p && !b->ta.refderivs.p)) {
    GRMP_TRY(upload_tables(tab_a, b->nq, su->grid->dim, s, &b->ta));
    b->sa = su; b->opa = op_a;
  }
  const i64 ncells = su->grid->ncells;
  GRMP_TRY(b->acoeffs.upload(coeffs_host, (size_t)su->ndofs, s));
  // the current iterate at the quadrature points: a_operator(u) and ansatz_operator(u) (nonlinearform.jl:124-141)
  for (int which = 0; which < 2; which++) {
    IiLocalParams ip{};
    ip.g = su->grid->view();
    GRMP_TRY(make_evalview(su, which == 0 ? op_a : b->op1, which == 0 ? b->ta : b->t1, &ip.e));
    ip.reg.n = 0; ip.nq = b->nq; ip.w = b->w.p; ip.kind = GRMP_II_NONE; ip.ardim = ip.e.rd; ip.factor = 1.0;
    DevBuf<double>& tab = which == 0 ? b->aq : b->gq;
    const size_t nt = (size_t)std::max<i64>(ncells * b->nq * ip.e.rd, 1);
    if (tab.n < nt) GRMP_TRY(tab.alloc(nt));
    if (b->aq_scratch.n < (size_t)std::max<i64>(ncells * ip.e.rd, 1)) GRMP_TRY(b->aq_scratch.alloc((size_t)std::max<i64>(ncells * ip.e.rd, 1)));
    ip.coeffs = b->acoeffs.p; ip.itemval = b->aq_scratch.p; ip.qtable = tab.p;
    GRMP_TRY(launch_ii_local(ip, s));
    if (which == 0) b->aq_rd = ip.e.rd;
  }
  EvalView e1, e2;
  GRMP_TRY(make_evalview(b->s1, b->op1, b->t1, &e1));
  GRMP_TRY(make_evalview(b->s2, b->op2, b->same_eval ? b->t1 : b->t2, &e2));
  if (b->aq_rd < 1 || e1.rd % b->aq_rd || e1.rd / b->aq_rd != e2.rd || e1.rd > 9)
    return fail(GRMP_EINVAL, "Newton convection: operator lengths do not fit (ansatz = ncomponents x xdim, a_operator = xdim, test = ncomponents)");
  GRMP_CUDA(cudaStreamSynchronize(s));
  if (!keep_pattern) b->have_pattern = false;
  b->have_values = false;
  return GRMP_OK;
}

int grmp_blf_newton_rhs(grmp_blf* b, double* b_host, int64_t offset) {
  if (!b || !b_host) return fail(GRMP_EINVAL, "grmp_blf_newton_rhs: NULL argument");
  if (b->action != GRMP_ACT_NEWTON_CONVECTION) return fail(GRMP_EINVAL, "not a Newton form");
  if (!b->have_values || !b->rbuf.p) return fail(GRMP_ESTATE, "grmp_blf_newton_rhs needs a prior grmp_blf_numeric");
  grmp_space* sp = b->s2;
  grmp_ctx* ctx = sp->grid->ctx;
  cudaStream_t s = ctx->stream;
  GRMP_CUDA(cudaSetDevice(ctx->device));
  if (!b->nl_dg_built) {
    GRMP_TRY(build_dofgather(s, sp->celldofs.p, sp->grid->ncells, sp->nd, sp->ndofs, &b->nl_dg));
    b->nl_dg_built = true;
  }
  GRMP_TRY(b->rhs.upload(b_host + offset, (size_t)sp->ndofs, s));
  GRMP_TRY(launch_lf_gather(s, b->nl_dg, b->rbuf.p, b->nl_active.p, b->rhs.p));
  GRMP_CUDA(cudaMemcpyAsync(b_host + offset, b->rhs.p, (size_t)sp->ndofs * 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  return GRMP_OK;
}

int grmp_blf_set_fixed_argument(grmp_blf* b, grmp_space* sa, int op_a, const grmp_evaltab* tab_a, const double* coeffs_host, int keep_pattern) {
  if (!b || !sa || !tab_a || !coeffs_host) return fail(GRMP_EINVAL, "grmp_blf_set_fixed_argument: NULL argument");
  if (b->action != GRMP_ACT_CONVECTION) return fail(GRMP_EINVAL, "the form was not created with GRMP_ACT_CONVECTION");
  if (sa->grid != b->s1->grid) return fail(GRMP_EINVAL, "spaces live on different grids");
  grmp_ctx* ctx = sa->grid->ctx;
  cudaStream_t s = ctx->stream;
  GRMP_CUDA(cudaSetDevice(ctx->device));
  if (b->sa != sa || b->opa != op_a || (!b->ta.refvals.p && !b->ta.refderivs.p)) {
    GRMP_TRY(upload_tables(tab_a, b->nq, sa->grid->dim, s, &b->ta));
    b->sa = sa; b->opa = op_a;
  }
  // a(x_q) = operator evaluation of the coefficient function at the quadrature points of every cell
  IiLocalParams ip{};
  ip.g = sa->grid->view();
  GRMP_TRY(make_evalview(sa, op_a, b->ta, &ip.e));
  EvalView e1;
  GRMP_TRY(make_evalview(b->s1, b->op1, b->t1, &e1));
  EvalView e2 = e1;
  if (!b->same_eval) GRMP_TRY(make_evalview(b->s2, b->op2, b->t2, &e2));
  if (ip.e.rd < 1 || e1.rd % ip.e.rd || e1.rd / ip.e.rd != e2.rd)
    return fail(GRMP_EINVAL, "convection: operator lengths do not fit (ansatz = ncomponents x xdim, coefficient = xdim, test = ncomponents)");
  const i64 ncells = sa->grid->ncells;
  b->aq_rd = ip.e.rd;
  ip.reg.n = 0; ip.nq = b->nq; ip.w = b->w.p; ip.kind = GRMP_II_NONE; ip.ardim = ip.e.rd; ip.factor = 1.0;
  GRMP_TRY(b->acoeffs.upload(coeffs_host, (size_t)sa->ndofs, s));
  const size_t nt = (size_t)std::max<i64>(ncells * b->nq * ip.e.rd, 1);
  if (b->aq.n < nt) GRMP_TRY(b->aq.alloc(nt));
  if (b->aq_scratch.n < (size_t)std::max<i64>(ncells * ip.e.rd, 1)) GRMP_TRY(b->aq_scratch.alloc((size_t)std::max<i64>(ncells * ip.e.rd, 1)));
  ip.coeffs = b->acoeffs.p; ip.data = nullptr; ip.b = nullptr; ip.itemval = b->aq_scratch.p; ip.qtable = b->aq.p;
  GRMP_TRY(launch_ii_local(ip, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  if (!keep_pattern) b->have_pattern = false;
  b->have_values = false;
  return GRMP_OK;
}

int grmp_blf_set_path(grmp_blf* b, int path) {
  if (!b || path < 0 || path > GRMP_PATH_COLOURED) return fail(GRMP_EINVAL, "grmp_blf_set_path: bad argument");
  b->path_req = path;
  return GRMP_OK;
}

int grmp_blf_symbolic(grmp_blf* b, double factor, int64_t* nnz_out) {
  if (!b) return fail(GRMP_EINVAL, "blf == NULL");
  grmp_ctx* ctx = b->s1->grid->ctx;
  cudaStream_t s = ctx->stream;
  GRMP_CUDA(cudaSetDevice(ctx->device));
  BlfLocalParams p;
  GRMP_TRY(fill_blf_params(b, factor, &p));
  const i64 ncells = p.g.ncells;
  const int nloc = p.e1.nd * p.e2.nd;
  const int req = b->path_req;
  const bool cellgrid = (p.g.xdim == p.g.dim);   // boundary-face grids: generic path (their items are embedded, the column kernels' geometry is not)
  const bool p2_ok = cellgrid && fast_p2tet_applicable(p);
  const bool col_ok = cellgrid && colpath_applicable(p, b->nq, &b->colp);
  const bool cell_req = (req == GRMP_PATH_ATOMIC || req == GRMP_PATH_COLOURED);
  if (req == GRMP_PATH_FAST && !p2_ok && !col_ok) return fail(GRMP_EUNSUPPORTED, "no fast path for this form");
  if ((req == GRMP_PATH_COLUMNS || cell_req) && !col_ok) return fail(GRMP_EUNSUPPORTED, "no column / cell kernel for this form");
  GRMP_CUDA(cudaEventRecord(ctx->ev0, s));
  DevBuf<u64> keys;
  GRMP_TRY(keys.alloc((size_t)ncells * nloc));
  p.keys = keys.p;
  GRMP_TRY(launch_blf_local(p, s));
  GRMP_TRY(build_pattern(s, keys, ncells * nloc, b->out_rows(), b->out_cols(), ncells, p.e1.nd, p.e2.nd, b->apt == GRMP_APT_SYMMETRIC,
                         cell_req /* per-cell local->nnz map: only the cell-parallel kernels scatter through it */, &b->pat));
  GRMP_TRY(b->nzval.alloc(b->pat.nnz));
  b->path = GRMP_PATH_GENERIC;
  b->p2tet_kappa = 0.0;
  bool want_p2 = p2_ok && (req == GRMP_PATH_AUTO || req == GRMP_PATH_FAST);
  if (want_p2 && req == GRMP_PATH_AUTO) {
    // cancellation guard: the ring-walk kernel derives diagonal entries from zero row sums, which loses digits on cells with
    // large sum_b |S_ab| / S_aa (slivers, strongly obtuse cells).  AUTO only takes it where 1e-12 is safe.
    GRMP_TRY(fast_p2tet_quality(ctx, p, &b->p2tet_kappa));
    const double kmax = getenv("GRMP_P2TET_KAPPA") ? atof(getenv("GRMP_P2TET_KAPPA")) : 16.0;
    if (!(b->p2tet_kappa <= kmax)) want_p2 = false;
  }
  if (want_p2) {
    const int rc = fast_p2tet_build(ctx, p, b->pat, b->w_host, b->t1_derivs_host, b->ncols_owned, b->s1->grid->geom_version, &b->fast);
    if (rc == GRMP_OK) b->path = GRMP_PATH_FAST;
    else if (rc != GRMP_EUNSUPPORTED || (req == GRMP_PATH_FAST && !col_ok)) return rc;   // grids the ring walk cannot order fall through
  }
  if (b->path == GRMP_PATH_GENERIC && col_ok && req != GRMP_PATH_GENERIC) {
    const int rc = colpath_build(ctx, p, b->pat, b->w_host, b->t1_vals_host, b->t1_derivs_host, b->t2_vals_host, b->t2_derivs_host,
                                 cell_req ? -1 : b->ncols_owned, cell_req, &b->colp);
    if (rc == GRMP_OK) {
      b->path = GRMP_PATH_COLUMNS;
      if (cell_req) {
        GRMP_TRY(cellpath_build(ctx, p, b->pat, req == GRMP_PATH_COLOURED, &b->colp));
        b->path = req;
      }
    } else if (rc != GRMP_EUNSUPPORTED || req != GRMP_PATH_AUTO) return rc;
  }
  b->pat.slotmap.release();
  GRMP_CUDA(cudaEventRecord(ctx->ev1, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  b->st.last_symbolic_ms = ms; b->st.nnz = b->pat.nnz; b->st.ncontrib = b->pat.ncontrib; b->st.path = b->path;
  b->st.ntiles = b->path == GRMP_PATH_FAST ? b->fast.ntiles : (int)b->colp.ntiles;
  b->have_pattern = true; b->have_values = false;
  b->csr.built = false;
  b->topo_grid = b->s1->grid->topo_version; b->topo_s1 = b->s1->topo_version; b->topo_s2 = b->s2->topo_version;
  b->halo_first_slot = -1;
  if (b->path == GRMP_PATH_FAST && b->ncols_owned >= 0 && b->ncols_owned < b->pat.ncols) {
    i64 cpv = 0;
    GRMP_CUDA(cudaMemcpy(&cpv, b->pat.colptr.p + b->ncols_owned, 8, cudaMemcpyDeviceToHost));
    b->halo_first_slot = cpv - 1;
    if (b->halo_first_slot < b->pat.nnz) {
      GRMP_CUDA(cudaMemsetAsync(b->nzval.p + b->halo_first_slot, 0, (size_t)(b->pat.nnz - b->halo_first_slot) * 8, s));
      GRMP_CUDA(cudaStreamSynchronize(s));
    }
  }
  if (nnz_out) *nnz_out = b->pat.nnz;
  return GRMP_OK;
}

int grmp_blf_get_pattern(grmp_blf* b, int64_t* colptr, int64_t* rowval) {
  if (!b || !colptr || (!rowval && b->pat.nnz)) return fail(GRMP_EINVAL, "grmp_blf_get_pattern: NULL argument");
  if (!b->have_pattern) return fail(GRMP_ESTATE, "grmp_blf_symbolic has not been called");
  cudaStream_t s = b->s1->grid->ctx->stream;
  GRMP_CUDA(cudaMemcpyAsync(colptr, b->pat.colptr.p, (size_t)(b->pat.ncols + 1) * 8, cudaMemcpyDeviceToHost, s));
  if (b->pat.nnz) GRMP_CUDA(cudaMemcpyAsync(rowval, b->pat.rowval.p, (size_t)b->pat.nnz * 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  return GRMP_OK;
}

int grmp_blf_numeric(grmp_blf* b, double factor, double* nzval_host) {
  if (!b) return fail(GRMP_EINVAL, "blf == NULL");
  GRMP_TRY(blf_check_fresh(b));
  grmp_ctx* ctx = b->s1->grid->ctx;
  cudaStream_t s = ctx->stream;
  GRMP_CUDA(cudaSetDevice(ctx->device));
  BlfLocalParams p;
  GRMP_TRY(fill_blf_params(b, factor, &p));
  GRMP_CUDA(cudaEventRecord(ctx->ev0, s));
  GRMP_TRY(blf_numeric_launch(b, p, s));
  GRMP_CUDA(cudaEventRecord(ctx->ev1, s));
  if (nzval_host && b->pat.nnz) GRMP_CUDA(cudaMemcpyAsync(nzval_host, b->nzval.p, (size_t)b->pat.nnz * 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  b->st.last_numeric_ms = ms;
  b->have_values = true;
  return GRMP_OK;
}

int grmp_blf_assemble_host(grmp_blf* b, double factor, const double* coords, const double* cellvolumes, const int32_t* cellnodes,
                           const int32_t* celldofs_row, const int32_t* celldofs_col, double* nzval_host) {
  if (!b || !coords || !cellvolumes) return fail(GRMP_EINVAL, "grmp_blf_assemble_host: NULL argument");
  GRMP_TRY(blf_check_fresh(b));
  grmp_grid* g = b->s1->grid;
  grmp_ctx* ctx = g->ctx;
  cudaStream_t s = ctx->stream, sc = ctx->copy_stream;
  GRMP_CUDA(cudaSetDevice(ctx->device));
  // geometry is what may change on a frozen pattern: it goes first on the compute stream.  The topology arrays, when the
  // caller hands them over, travel on the copy stream concurrently with the kernels and the download (PCIe is full duplex)
  // and are compared with the arrays of the symbolic pass.
  GRMP_TRY(g->coords.upload(coords, (size_t)g->nnodes * g->xdim, s));
  GRMP_TRY(g->vol.upload(cellvolumes, (size_t)g->ncells, s));
  g->geom_version++;
  const bool check = cellnodes || celldofs_row || celldofs_col;
  if (check) {
    if (b->cmp_flag.n == 0) GRMP_TRY(b->cmp_flag.alloc(1));
    GRMP_CUDA(cudaMemsetAsync(b->cmp_flag.p, 0, sizeof(int), sc));
    auto cmp = [&](const int32_t* host, DevBuf<i32>& scratch, const DevBuf<i32>& ref) -> int {
      if (!host || ref.n == 0) return GRMP_OK;
      GRMP_TRY(scratch.upload(host, ref.n, sc));
      compare_i32<<<(unsigned)((ref.n + 255) / 256), 256, 0, sc>>>(scratch.p, ref.p, (i64)ref.n, b->cmp_flag.p);
      GRMP_CUDA(cudaGetLastError());
      return GRMP_OK;
    };
    GRMP_TRY(cmp(cellnodes, b->cmp_nodes, g->cellnodes));
    GRMP_TRY(cmp(celldofs_row, b->cmp_dofs1, b->s1->celldofs));
    if (b->s2 != b->s1) GRMP_TRY(cmp(celldofs_col, b->cmp_dofs2, b->s2->celldofs));
  }
  BlfLocalParams p;
  GRMP_TRY(fill_blf_params(b, factor, &p));
  GRMP_CUDA(cudaEventRecord(ctx->ev0, s));
  GRMP_TRY(blf_numeric_launch(b, p, s));
  GRMP_CUDA(cudaEventRecord(ctx->ev1, s));
  if (nzval_host && b->pat.nnz) GRMP_CUDA(cudaMemcpyAsync(nzval_host, b->nzval.p, (size_t)b->pat.nnz * 8, cudaMemcpyDeviceToHost, s));
  int differs = 0;
  if (check) GRMP_CUDA(cudaMemcpyAsync(&differs, b->cmp_flag.p, sizeof(int), cudaMemcpyDeviceToHost, sc));
  GRMP_CUDA(cudaStreamSynchronize(s));
  GRMP_CUDA(cudaStreamSynchronize(sc));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  b->st.last_numeric_ms = ms;
  b->have_values = true;
  if (differs) {
    b->have_values = false;
    return fail(GRMP_ESTATE, "grmp_blf_assemble_host: CellNodes / CellDofs differ from the arrays of the symbolic pass (stale pattern)");
  }
  return GRMP_OK;
}

int grmp_blf_numeric_steps(grmp_blf* b, double factor, int nsteps, double* total_ms) {
  if (!b || nsteps < 0) return fail(GRMP_EINVAL, "grmp_blf_numeric_steps: bad argument");
  GRMP_TRY(blf_check_fresh(b));
  grmp_ctx* ctx = b->s1->grid->ctx;
  cudaStream_t s = ctx->stream;
  GRMP_CUDA(cudaSetDevice(ctx->device));
  BlfLocalParams p;
  GRMP_TRY(fill_blf_params(b, factor, &p));
  // Steps are replayed as a CUDA graph holding a chunk of steps: smaller gaps between the launches, and the programmatic edges
  // between the kernels of consecutive steps (fast path) survive inside the graph.  The graph is set-up work like the symbolic
  // pass: captured and instantiated before the timed region starts (capturing launches nothing).
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int chunk = 0;
  if (nsteps > 2 && !getenv("GRMP_NO_GRAPH")) {
    chunk = std::min(nsteps, getenv("GRMP_GRAPH_STEPS") ? std::max(1, atoi(getenv("GRMP_GRAPH_STEPS"))) : 8);
    GRMP_TRY(blf_numeric_launch(b, p, s));      // one plain step first: lazy one-time work (coordinate repack, attributes) stays out of the capture
    GRMP_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    int rc = GRMP_OK;
    for (int j = 0; j < chunk && rc == GRMP_OK; j++) rc = blf_numeric_launch(b, p, s);
    const cudaError_t ce = cudaStreamEndCapture(s, &graph);
    if (!(rc == GRMP_OK && ce == cudaSuccess && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess)) {
      (void)cudaGetLastError();     // capture not possible: plain launches below
      if (exec) cudaGraphExecDestroy(exec);
      exec = nullptr;
    } else {
      cudaGraphUpload(exec, s);
    }
  }
  GRMP_CUDA(cudaStreamSynchronize(s));
  GRMP_CUDA(cudaEventRecord(ctx->ev0, s));
  int k = 0;
  if (exec)
    for (; k + chunk <= nsteps; k += chunk) GRMP_CUDA(cudaGraphLaunch(exec, s));
  for (; k < nsteps; k++) GRMP_TRY(blf_numeric_launch(b, p, s));
  GRMP_CUDA(cudaEventRecord(ctx->ev1, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  if (exec) cudaGraphExecDestroy(exec);
  if (graph) cudaGraphDestroy(graph);
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  if (total_ms) *total_ms = ms;
  if (nsteps > 0) { b->st.last_numeric_ms = ms / nsteps; b->have_values = true; }
  if (b->path == GRMP_PATH_FAST) GRMP_TRY(fast_p2tet_print_prof(ctx, b->fast));
  return GRMP_OK;
}

int grmp_blf_set_owned_columns(grmp_blf* b, int64_t ncols_owned) {
  if (!b) return fail(GRMP_EINVAL, "blf == NULL");
  b->ncols_owned = ncols_owned;
  return GRMP_OK;
}

int grmp_blf_get_values(grmp_blf* b, double* nzval_host) {
  if (!b || (!nzval_host && b->pat.nnz)) return fail(GRMP_EINVAL, "grmp_blf_get_values: NULL argument");
  if (!b->have_values) return fail(GRMP_ESTATE, "grmp_blf_numeric has not been called");
  cudaStream_t s = b->s1->grid->ctx->stream;
  if (b->pat.nnz) GRMP_CUDA(cudaMemcpyAsync(nzval_host, b->nzval.p, (size_t)b->pat.nnz * 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  return GRMP_OK;
}

int grmp_blf_transpose_copy(grmp_blf* b, double factor, double factor_transpose, int64_t* colptr_t, int64_t* rowval_t, double* nzval_t) {
  if (!b || !colptr_t) return fail(GRMP_EINVAL, "grmp_blf_transpose_copy: NULL argument");
  if (!b->have_values) return fail(GRMP_ESTATE, "transpose copy needs a prior grmp_blf_numeric with the same factor");
  if (b->apt == GRMP_APT_SYMMETRIC) return fail(GRMP_EUNSUPPORTED, "transpose_copy is only assembled by the general branch (bilinearform.jl:347-367)");
  cudaStream_t s = b->s1->grid->ctx->stream;
  DevBuf<i64> cpt, rvt; DevBuf<i32> perm; DevBuf<double> tv, tvp;
  GRMP_TRY(build_transposed(s, b->pat, cpt, rvt, perm));
  GRMP_TRY(tv.alloc(b->pat.nnz)); GRMP_TRY(tvp.alloc(b->pat.nnz));
  if (b->path == GRMP_PATH_GENERIC && b->lbuf.n != 0) {
    GRMP_TRY(launch_gather_transposed(s, b->pat, b->lbuf.p, factor, factor_transpose, tv.p));     // term by term, bit-exact
  } else {
    // the fast kernels keep no per-cell values: scale the assembled entries (equal to the term-by-term sum up to rounding)
    GRMP_TRY(launch_scale(s, b->nzval.p, b->pat.nnz, -(factor_transpose / factor), tv.p));
  }
  GRMP_TRY(launch_permute(s, tv.p, perm.p, b->pat.nnz, tvp.p));
  GRMP_CUDA(cudaMemcpyAsync(colptr_t, cpt.p, (size_t)(b->pat.nrows + 1) * 8, cudaMemcpyDeviceToHost, s));
  if (b->pat.nnz) {
    GRMP_CUDA(cudaMemcpyAsync(rowval_t, rvt.p, (size_t)b->pat.nnz * 8, cudaMemcpyDeviceToHost, s));
    GRMP_CUDA(cudaMemcpyAsync(nzval_t, tvp.p, (size_t)b->pat.nnz * 8, cudaMemcpyDeviceToHost, s));
  }
  GRMP_CUDA(cudaStreamSynchronize(s));
  return GRMP_OK;
}

int grmp_blf_stats(grmp_blf* b, grmp_stats* out) {
  if (!b || !out) return fail(GRMP_EINVAL, "NULL argument");
  *out = b->st;
  return GRMP_OK;
}

int grmp_blf_device_values(grmp_blf* b, void** dptr) {
  if (!b || !dptr) return fail(GRMP_EINVAL, "NULL argument");
  *dptr = b->nzval.p;
  return GRMP_OK;
}

int grmp_blf_device_csc(grmp_blf* b, grmp_device_csc* out) {
  if (!b || !out) return fail(GRMP_EINVAL, "NULL argument");
  if (!b->have_values) return fail(GRMP_ESTATE, "grmp_blf_numeric has not been called");
  out->nrows = b->pat.nrows; out->ncols = b->pat.ncols; out->nnz = b->pat.nnz;
  out->colptr = b->pat.colptr.p; out->rowval = b->pat.rowval.p; out->nzval = b->nzval.p;
  out->device = b->s1->grid->ctx->device; out->reserved = 0;
  return GRMP_OK;
}

int grmp_blf_matmul_device(grmp_blf* b, const double* b_dev, double* a_dev, double factor, int transposed) {
  if (!b || !b_dev || !a_dev) return fail(GRMP_EINVAL, "grmp_blf_matmul_device: NULL argument");
  if (!b->have_values) return fail(GRMP_ESTATE, "grmp_blf_numeric has not been called");
  grmp_ctx* ctx = b->s1->grid->ctx;
  GRMP_CUDA(cudaSetDevice(ctx->device));
  if (!transposed && !b->csr.built) GRMP_TRY(build_csr_view(ctx->stream, b->pat, &b->csr));
  GRMP_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  GRMP_TRY(launch_matmul(ctx->stream, b->pat, b->csr, b->nzval.p, b_dev, a_dev, factor, transposed));
  GRMP_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  GRMP_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  b->last_matmul_ms = ms;
  return GRMP_OK;
}

int grmp_blf_matmul(grmp_blf* b, const double* b_host, double* a_host, double factor, int transposed) {
  if (!b || !b_host || !a_host) return fail(GRMP_EINVAL, "grmp_blf_matmul: NULL argument");
  if (!b->have_values) return fail(GRMP_ESTATE, "grmp_blf_numeric has not been called");
  cudaStream_t s = b->s1->grid->ctx->stream;
  const i64 nin = transposed ? b->pat.nrows : b->pat.ncols, nout = transposed ? b->pat.ncols : b->pat.nrows;
  GRMP_TRY(b->vec_in.upload(b_host, (size_t)nin, s));
  GRMP_TRY(b->vec_out.upload(a_host, (size_t)nout, s));
  GRMP_TRY(grmp_blf_matmul_device(b, b->vec_in.p, b->vec_out.p, factor, transposed));
  if (nout) GRMP_CUDA(cudaMemcpyAsync(a_host, b->vec_out.p, (size_t)nout * 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  return GRMP_OK;
}

int grmp_blf_residual(grmp_blf* b, const double* x_host, const double* b_host, const int64_t* fixed_dofs, int64_t nfixed, double* r_host,
                      double* norm2) {
  if (!b || !x_host || nfixed < 0 || (nfixed > 0 && !fixed_dofs)) return fail(GRMP_EINVAL, "grmp_blf_residual: bad argument");
  if (!b->have_values) return fail(GRMP_ESTATE, "grmp_blf_numeric has not been called");
  grmp_ctx* ctx = b->s1->grid->ctx;
  cudaStream_t s = ctx->stream;
  GRMP_CUDA(cudaSetDevice(ctx->device));
  const i64 n = b->pat.nrows;
  GRMP_TRY(b->vec_in.upload(x_host, (size_t)b->pat.ncols, s));
  if (b->vec_out.n != (size_t)n) GRMP_TRY(b->vec_out.alloc(n));
  if (n) GRMP_CUDA(cudaMemsetAsync(b->vec_out.p, 0, (size_t)n * 8, s));
  if (!b->csr.built) GRMP_TRY(build_csr_view(s, b->pat, &b->csr));
  GRMP_TRY(launch_matmul(s, b->pat, b->csr, b->nzval.p, b->vec_in.p, b->vec_out.p, 1.0, 0));
  if (b_host) GRMP_TRY(b->vec_rhs.upload(b_host, (size_t)n, s));
  if (nfixed > 0) GRMP_TRY(b->fixed.upload(fixed_dofs, (size_t)nfixed, s));
  if (b->scalar.n == 0) GRMP_TRY(b->scalar.alloc(2));
  GRMP_TRY(launch_residual_finish(s, b->vec_out.p, b_host ? b->vec_rhs.p : nullptr, n, nfixed > 0 ? b->fixed.p : nullptr, nfixed,
                                  norm2 ? b->scalar.p : nullptr));
  if (norm2) GRMP_CUDA(cudaMemcpyAsync(norm2, b->scalar.p, 8, cudaMemcpyDeviceToHost, s));
  if (r_host && n) GRMP_CUDA(cudaMemcpyAsync(r_host, b->vec_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  return GRMP_OK;
}

int grmp_blf_apply_penalties(grmp_blf* b, const int64_t* fixed_dofs, int64_t nfixed, double penalty, int64_t* nmissing) {
  if (!b || nfixed < 0 || (nfixed > 0 && !fixed_dofs)) return fail(GRMP_EINVAL, "grmp_blf_apply_penalties: bad argument");
  if (!b->have_values) return fail(GRMP_ESTATE, "grmp_blf_numeric has not been called");
  grmp_ctx* ctx = b->s1->grid->ctx;
  cudaStream_t s = ctx->stream;
  GRMP_CUDA(cudaSetDevice(ctx->device));
  if (nmissing) *nmissing = 0;
  if (nfixed == 0) return GRMP_OK;
  GRMP_TRY(b->fixed.upload(fixed_dofs, (size_t)nfixed, s));
  DevBuf<i64> missing;
  GRMP_TRY(missing.alloc(1));
  GRMP_TRY(launch_penalties(s, b->pat, b->nzval.p, b->fixed.p, nfixed, penalty, missing.p));
  i64 m = 0;
  GRMP_CUDA(cudaMemcpyAsync(&m, missing.p, 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  if (nmissing) *nmissing = m;
  if (m > 0) return fail(GRMP_EUNSUPPORTED, "apply_penalties: a fixed dof has no stored diagonal entry (the frozen pattern cannot grow)");
  return GRMP_OK;
}

int grmp_lf_create(grmp_space* sp, int op, const int32_t* regions, int nregions, int nq, const double* qweights, const grmp_evaltab* tab,
                   grmp_lf** out) {
  if (!sp || !qweights || !tab || !out || nq <= 0) return fail(GRMP_EINVAL, "grmp_lf_create: bad argument");
  grmp_lf* l = new grmp_lf();
  cudaStream_t s = sp->grid->ctx->stream;
  l->sp = sp; l->op = op; l->nq = nq;
  int rc = make_regions(regions, nregions, &l->reg);
  if (!rc) rc = l->w.upload(qweights, nq, s);
  if (!rc) rc = upload_tables(tab, nq, sp->grid->dim, s, &l->tab);
  l->w_host.assign(qweights, qweights + nq);
  if (!rc && tab->refvals) l->vals_host.assign(tab->refvals, tab->refvals + (size_t)nq * tab->nd_all * tab->ncomp);
  if (!rc && tab->refderivs) l->derivs_host.assign(tab->refderivs, tab->refderivs + (size_t)nq * sp->grid->dim * tab->nd_all * tab->ncomp);
  l->topo_grid = sp->grid->topo_version; l->topo_sp = sp->topo_version;
  EvalView e;
  if (!rc) rc = make_evalview(sp, op, l->tab, &e);
  if (!rc) rc = build_dofgather(s, sp->celldofs.p, sp->grid->ncells, sp->nd, sp->ndofs, &l->dg);
  if (!rc) rc = l->lbuf.alloc((size_t)sp->nd * sp->grid->ncells);
  if (!rc) rc = l->active.alloc((size_t)sp->grid->ncells);
  if (!rc) rc = l->b.alloc((size_t)sp->ndofs);
  if (!rc && cudaStreamSynchronize(s) != cudaSuccess) rc = fail(GRMP_ECUDA, "upload failed");
  if (rc) { delete l; return rc; }
  *out = l;
  return GRMP_OK;
}
int grmp_lf_destroy(grmp_lf* l) { delete l; return GRMP_OK; }

int grmp_lf_set_path(grmp_lf* l, int path) {
  if (!l || (path != GRMP_PATH_AUTO && path != GRMP_PATH_GENERIC && path != GRMP_PATH_FAST && path != GRMP_PATH_COLUMNS))
    return fail(GRMP_EINVAL, "grmp_lf_set_path: bad argument");
  l->path_req = path;
  l->fast_tried = false;
  return GRMP_OK;
}

static int lf_assemble_impl(grmp_lf* l, double factor, int fsrc, const double* fdata, bool fdata_resident, double* b_host, int64_t offset) {
  if (!l || !b_host) return fail(GRMP_EINVAL, "grmp_lf_assemble: NULL argument");
  if (fsrc != GRMP_F_NONE && !fdata && !fdata_resident) return fail(GRMP_EINVAL, "fdata missing");
  grmp_space* sp = l->sp;
  grmp_ctx* ctx = sp->grid->ctx;
  cudaStream_t s = ctx->stream;
  GRMP_CUDA(cudaSetDevice(ctx->device));
  if (l->topo_grid != sp->grid->topo_version || l->topo_sp != sp->topo_version)
    return fail(GRMP_ESTATE, "CellNodes / CellDofs changed after grmp_lf_create: the gather lists are stale, create the linear form again");
  LfLocalParams p{};
  p.g = sp->grid->view();
  GRMP_TRY(make_evalview(sp, l->op, l->tab, &p.e));
  p.reg = l->reg; p.nq = l->nq; p.w = l->w.p; p.factor = factor; p.fsrc = fsrc;
  if (!l->fast_tried) {      // owner-computes gather kernels where one exists for the evaluator (AUTO), else the bit-exact two-phase path
    l->fast_tried = true;
    l->path = GRMP_PATH_GENERIC;
    if (l->path_req != GRMP_PATH_GENERIC) {
      int rc = GRMP_EUNSUPPORTED;
      if (sp->grid->xdim == sp->grid->dim && lfpath_applicable(p.e, sp->grid->dim, l->nq, &l->lfp))
        rc = lfpath_build(ctx, p.g, p.e, l->reg, l->w_host, l->vals_host, l->derivs_host, sp->ndofs, &l->lfp);
      else set_error("no linear form kernel for this evaluator");
      if (rc == GRMP_OK) l->path = GRMP_PATH_COLUMNS;
      else if (rc != GRMP_EUNSUPPORTED || l->path_req != GRMP_PATH_AUTO) { l->fast_tried = false; return rc; }
    }
  }
  if (fdata_resident) { /* l->fdata already holds the table (grmp_lf_assemble_feb) */ }
  else if (fsrc == GRMP_F_CONST) GRMP_TRY(l->fdata.upload(fdata, (size_t)p.e.rd, s));
  else if (fsrc == GRMP_F_QP_TABLE) GRMP_TRY(l->fdata.upload(fdata, (size_t)sp->grid->ncells * l->nq * p.e.rd, s));
  p.fdata = l->fdata.p; p.lbuf = l->lbuf.p; p.active = l->active.p;
  GRMP_CUDA(cudaMemcpyAsync(l->b.p, b_host + offset, (size_t)sp->ndofs * 8, cudaMemcpyHostToDevice, s));
  GRMP_CUDA(cudaEventRecord(ctx->ev0, s));
  if (l->path == GRMP_PATH_COLUMNS) {
    GRMP_TRY(lfpath_numeric(ctx, p.g, l->lfp, factor, fsrc, l->fdata.p, l->b.p));
  } else {
    GRMP_TRY(launch_lf_local(p, s));
    GRMP_TRY(launch_lf_gather(s, l->dg, l->lbuf.p, l->active.p, l->b.p));
  }
  GRMP_CUDA(cudaEventRecord(ctx->ev1, s));
  GRMP_CUDA(cudaMemcpyAsync(b_host + offset, l->b.p, (size_t)sp->ndofs * 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  l->st.last_numeric_ms = ms; l->st.kernel_launches = l->path == GRMP_PATH_COLUMNS ? (i64)l->lfp.classes.size() : 2; l->st.path = l->path;
  return GRMP_OK;
}

int grmp_ii_create(grmp_space* sp, int op, int kind, const int32_t* regions, int nregions, int nq, const double* qweights,
                   const grmp_evaltab* tab, grmp_ii** out) {
  if (!sp || !qweights || !tab || !out || nq <= 0 || kind < GRMP_II_NONE || kind > GRMP_II_L2ERROR) return fail(GRMP_EINVAL, "grmp_ii_create: bad argument");
  grmp_ii* ii = new grmp_ii();
  cudaStream_t s = sp->grid->ctx->stream;
  ii->sp = sp; ii->op = op; ii->kind = kind; ii->nq = nq;
  int rc = make_regions(regions, nregions, &ii->reg);
  if (!rc) rc = ii->w.upload(qweights, nq, s);
  if (!rc) rc = upload_tables(tab, nq, sp->grid->dim, s, &ii->tab);
  ii->topo_grid = sp->grid->topo_version; ii->topo_sp = sp->topo_version;
  EvalView e;
  if (!rc) rc = make_evalview(sp, op, ii->tab, &e);
  if (!rc && cudaStreamSynchronize(s) != cudaSuccess) rc = fail(GRMP_ECUDA, "upload failed");
  if (rc) { delete ii; return rc; }
  *out = ii;
  return GRMP_OK;
}
int grmp_ii_destroy(grmp_ii* ii) { delete ii; return GRMP_OK; }

int grmp_ii_resultdim(grmp_ii* ii, int* resultdim) {
  if (!ii || !resultdim) return fail(GRMP_EINVAL, "NULL argument");
  EvalView e;
  GRMP_TRY(make_evalview(ii->sp, ii->op, ii->tab, &e));
  *resultdim = ii->kind == GRMP_II_NONE ? e.rd : 1;
  return GRMP_OK;
}

int grmp_ii_evaluate(grmp_ii* ii, const double* coeffs_host, double factor, const double* data_host, double* b_host, double* total_host) {
  if (!ii || !coeffs_host) return fail(GRMP_EINVAL, "grmp_ii_evaluate: NULL argument");
  if (ii->kind == GRMP_II_L2ERROR && !data_host) return fail(GRMP_EINVAL, "grmp_ii_evaluate: compare data missing");
  grmp_space* sp = ii->sp;
  grmp_ctx* ctx = sp->grid->ctx;
  cudaStream_t s = ctx->stream;
  GRMP_CUDA(cudaSetDevice(ctx->device));
  if (ii->topo_grid != sp->grid->topo_version || ii->topo_sp != sp->topo_version)
    return fail(GRMP_ESTATE, "CellNodes / CellDofs changed after grmp_ii_create: create the integrator again");
  IiLocalParams p{};
  p.g = sp->grid->view();
  GRMP_TRY(make_evalview(sp, ii->op, ii->tab, &p.e));
  const i64 ncells = sp->grid->ncells;
  const int ardim = ii->kind == GRMP_II_NONE ? p.e.rd : 1;
  p.reg = ii->reg; p.nq = ii->nq; p.w = ii->w.p; p.kind = ii->kind; p.ardim = ardim; p.factor = factor;
  GRMP_TRY(ii->coeffs.upload(coeffs_host, (size_t)sp->ndofs, s));
  if (ii->kind == GRMP_II_L2ERROR) GRMP_TRY(ii->data.upload(data_host, (size_t)ncells * ii->nq * p.e.rd, s));
  if (b_host) GRMP_TRY(ii->b.upload(b_host, (size_t)ncells * ardim, s));
  if (ii->itemval.n < (size_t)ncells * ardim) GRMP_TRY(ii->itemval.alloc((size_t)std::max<i64>(ncells * ardim, 1)));
  if (ii->total.n < (size_t)ardim) GRMP_TRY(ii->total.alloc((size_t)ardim));
  p.coeffs = ii->coeffs.p; p.data = ii->data.p; p.b = b_host ? ii->b.p : nullptr; p.itemval = ii->itemval.p;
  GRMP_TRY(launch_ii_local(p, s));
  if (total_host) {
    for (int j = 0; j < ardim; j++) GRMP_TRY(device_sum(s, ii->itemval.p + (size_t)j * ncells, ncells, ii->total.p + j, &ii->tmp));
    GRMP_CUDA(cudaMemcpyAsync(total_host, ii->total.p, (size_t)ardim * 8, cudaMemcpyDeviceToHost, s));
  }
  if (b_host) GRMP_CUDA(cudaMemcpyAsync(b_host, ii->b.p, (size_t)ncells * ardim * 8, cudaMemcpyDeviceToHost, s));
  GRMP_CUDA(cudaStreamSynchronize(s));
  return GRMP_OK;
}

int grmp_lf_assemble(grmp_lf* l, double factor, int fsrc, const double* fdata, double* b_host, int64_t offset) {
  return lf_assemble_impl(l, factor, fsrc, fdata, false, b_host, offset);
}

int grmp_lf_assemble_feb(grmp_lf* l, double factor, grmp_space* sa, int op_a, const grmp_evaltab* tab_a, const double* coeffs_host,
                         double* b_host, int64_t offset) {
  if (!l || !sa || !tab_a || !coeffs_host || !b_host) return fail(GRMP_EINVAL, "grmp_lf_assemble_feb: NULL argument");
  grmp_space* sp = l->sp;
  if (sa->grid != sp->grid) return fail(GRMP_EINVAL, "spaces live on different grids");
  grmp_ctx* ctx = sp->grid->ctx;
  cudaStream_t s = ctx->stream;
  GRMP_CUDA(cudaSetDevice(ctx->device));
  if (l->sa != sa || l->opa != op_a || (!l->ta.refvals.p && !l->ta.refderivs.p)) {
    GRMP_TRY(upload_tables(tab_a, l->nq, sa->grid->dim, s, &l->ta));
    l->sa = sa; l->opa = op_a;
  }
  IiLocalParams ip{};
  ip.g = sa->grid->view();
  GRMP_TRY(make_evalview(sa, op_a, l->ta, &ip.e));
  EvalView et;
  GRMP_TRY(make_evalview(sp, l->op, l->tab, &et));
  if (ip.e.rd != et.rd) return fail(GRMP_EINVAL, "operator result lengths of the coefficient argument and the test function differ");
  const i64 ncells = sa->grid->ncells;
  ip.reg.n = 0; ip.nq = l->nq; ip.w = l->w.p; ip.kind = GRMP_II_NONE; ip.ardim = ip.e.rd; ip.factor = 1.0;
  GRMP_TRY(l->acoeffs.upload(coeffs_host, (size_t)sa->ndofs, s));
  const size_t nt = (size_t)std::max<i64>(ncells * l->nq * ip.e.rd, 1);
  if (l->fdata.n < nt) GRMP_TRY(l->fdata.alloc(nt));
  if (l->ascratch.n < (size_t)std::max<i64>(ncells * ip.e.rd, 1)) GRMP_TRY(l->ascratch.alloc((size_t)std::max<i64>(ncells * ip.e.rd, 1)));
  ip.coeffs = l->acoeffs.p; ip.itemval = l->ascratch.p; ip.qtable = l->fdata.p;
  GRMP_TRY(launch_ii_local(ip, s));
  return lf_assemble_impl(l, factor, GRMP_F_QP_TABLE, nullptr, true, b_host, offset);
}

int grmp_lf_stats(grmp_lf* l, grmp_stats* out) {
  if (!l || !out) return fail(GRMP_EINVAL, "NULL argument");
  *out = l->st;
  return GRMP_OK;
}

}  // extern "C"
