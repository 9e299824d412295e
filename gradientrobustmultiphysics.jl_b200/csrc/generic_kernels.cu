// generic_kernels.cu -- per-cell local matrices / vectors in the reference's operation order.
//
// Compiled with -fmad=false.  One thread evaluates one cell exactly like the reference's
// serial loop body (update_basis! variants + the quadrature/dof loops of assemble!), so
// that (i) the exact-zero test of _addnz (src/fematrix.jl:54-58), which decides the
// sparsity pattern, and (ii) the values handed to rawupdateindex! are bit-identical to an
// un-fused CPU evaluation.  The kernels here feed the symbolic pass (keys) and the generic
// numeric path (element-matrix buffer + ordered gather); the roofline kernel for the
// metric configuration lives in fastpath_p2tet.cu.
//
//   update_trafo!/mapderiv!            src/feevaluator.jl:371-390 (ExtendableGrids semantics)
//   update_basis! H1                   src/feevaluator_h1.jl:2-14, 33-43, 61-145
//   update_basis! ReconstructionId     src/feevaluator_h1.jl:342-381, src/reconstructions.jl:353-535
//   update_basis! Hdiv                 src/feevaluator_hdiv.jl:2-19, 54-71
//   BLF loops                          src/assemblypatterns/bilinearform.jl:294-368
//   LF loops                           src/assemblypatterns/linearform.jl:181-220
#include "common.cuh"

namespace grmp {

constexpr int RDMAX = 9;

struct Geo {
  double A[3][3], Ainv[3][3], det;
};

// update_trafo! : A[:,j] = x_{j+1} - x_1 ; det from A
__device__ __forceinline__ void geo_update(const GridView& g, i64 cell, Geo& T) {
  const int d = g.dim;
  if (g.xdim != d) { T.det = g.vol[cell]; return; }   // boundary-face items: Identity / NormalFlux evaluators only (make_evalview); no affine map, det carries |F|
  const i32* cn = g.cellnodes + cell * (d + 1);
  const double* x0 = g.coords + (i64)(cn[0] - 1) * d;
  double b[3];
  for (int k = 0; k < d; k++) b[k] = x0[k];
  for (int j = 0; j < d; j++) {
    const double* xj = g.coords + (i64)(cn[j + 1] - 1) * d;
    for (int k = 0; k < d; k++) T.A[k][j] = xj[k] - b[k];
  }
  if (d == 2)
    T.det = T.A[0][0] * T.A[1][1] - T.A[0][1] * T.A[1][0];
  else
    T.det = T.A[0][0] * (T.A[1][1] * T.A[2][2] - T.A[1][2] * T.A[2][1]) - T.A[0][1] * (T.A[1][0] * T.A[2][2] - T.A[1][2] * T.A[2][0]) +
            T.A[0][2] * (T.A[1][0] * T.A[2][1] - T.A[1][1] * T.A[2][0]);
}
// mapderiv! : Ainv = A^{-T}, det = d! * |T|
__device__ __forceinline__ void geo_mapderiv(const GridView& g, i64 cell, Geo& T) {
  const double(*A)[3] = T.A;
  if (g.dim == 2) {
    double dt = 2 * g.vol[cell];
    T.Ainv[1][1] = A[0][0] / dt;
    T.Ainv[1][0] = -A[0][1] / dt;
    T.Ainv[0][1] = -A[1][0] / dt;
    T.Ainv[0][0] = A[1][1] / dt;
  } else {
    double dt = 6 * g.vol[cell];
    T.Ainv[0][0] = (A[1][1] * A[2][2] - A[1][2] * A[2][1]) / dt;
    T.Ainv[0][1] = -(A[1][0] * A[2][2] - A[1][2] * A[2][0]) / dt;
    T.Ainv[0][2] = (A[1][0] * A[2][1] - A[1][1] * A[2][0]) / dt;
    T.Ainv[1][0] = -(A[0][1] * A[2][2] - A[0][2] * A[2][1]) / dt;
    T.Ainv[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) / dt;
    T.Ainv[1][2] = -(A[0][0] * A[2][1] - A[0][1] * A[2][0]) / dt;
    T.Ainv[2][0] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) / dt;
    T.Ainv[2][1] = -(A[0][0] * A[1][2] - A[0][2] * A[1][0]) / dt;
    T.Ainv[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) / dt;
  }
}

__device__ __forceinline__ bool cell_active(const GridView& g, const RegionFilter& r, i64 cell) {
  if (r.n == 0) return true;
  if (!g.regions) return false;
  i32 cr = g.regions[cell];
  for (int k = 0; k < r.n; k++)
    if (cr == r.r[k]) return true;
  return false;
}

// per-cell data of one evaluator: coefficients[k][dof] and the basis subset
template <int NDMAX> struct CellCoef {
  double co[3][NDMAX];
  int subset[NDMAX];
};

// get_coefficients / get_basissubset closures for family `fam` with `ndc` dofs on the cell
template <int NDMAX>
__device__ void cell_coefficients(const GridView& g, i64 cell, int fam, int ndc, int ncomp, CellCoef<NDMAX>& c) {
  const int d = g.dim, nf = d + 1;
  for (int dof = 0; dof < ndc; dof++) {
    c.subset[dof] = dof;
    for (int k = 0; k < ncomp; k++) c.co[k][dof] = 1.0;
  }
  if (g.xdim != d) return;   // face bases carry no coefficients (the boundary face is the first face of its only cell: sign +1)
  if (fam == FAM_H1BR) {  // h1v_br.jl:150-162, 253-273: bubble columns = face normal
    const i32* cf = g.cellfaces + cell * nf;
    for (int f = 0; f < nf; f++)
      for (int k = 0; k < d; k++) c.co[k][d * nf + f] = g.fnormals[(i64)(cf[f] - 1) * d + k];
  } else if (fam == FAM_RT0) {  // hdiv_rt0.jl:106-116
    const i32* sg = g.signs + cell * nf;
    for (int j = 0; j < nf; j++)
      for (int k = 0; k < ncomp; k++) c.co[k][j] = (double)sg[j];
  } else if (fam == FAM_BDM1 && d == 2) {  // hdiv_bdm1.jl (2D): sign on the RT0 functions only
    const i32* sg = g.signs + cell * nf;
    for (int j = 0; j < nf; j++)
      for (int k = 0; k < d; k++) c.co[k][2 * j] = (double)sg[j];
  } else if (fam == FAM_BDM1) {  // hdiv_bdm1.jl (3D): sign / -1 / +1, subset by face orientation
    const i32* sg = g.signs + cell * nf;
    const i32* o = g.orient + cell * nf;
    const int s1[4] = {1, 0, 1, 2}, s2[4] = {2, 2, 0, 1};
    for (int j = 0; j < nf; j++) {
      for (int k = 0; k < d; k++) {
        c.co[k][3 * j] = (double)sg[j];
        c.co[k][3 * j + 1] = -1.0;
        c.co[k][3 * j + 2] = 1.0;
      }
      c.subset[3 * j] = 4 * (j + 1) - 3 - 1;
      c.subset[3 * j + 1] = 4 * (j + 1) - s1[o[j] - 1] - 1;
      c.subset[3 * j + 2] = 4 * (j + 1) - s2[o[j] - 1] - 1;
    }
  }
}

// boundary_coefficients! BR -> RT0 / BDM1 (reconstructions.jl:353-403, 474-535); rc[dofBR][dofR]
template <int NDMAX>
__device__ void recon_coefficients(const GridView& g, i64 cell, int op, double (*rc)[12]) {
  const int d = g.dim, nf = d + 1;
  const i32* cf = g.cellfaces + cell * nf;
  const int TRI_FACE[3][2] = {{0, 1}, {1, 2}, {2, 0}};
  const int TET_FACE[4][3] = {{0, 2, 1}, {0, 1, 3}, {1, 2, 3}, {0, 3, 2}};
  if (d == 2) {
    for (int f = 0; f < 3; f++) {
      i64 face = cf[f] - 1;
      double fv = g.fvol[face];
      for (int n = 0; n < 2; n++) {
        int node = TRI_FACE[f][n];
        for (int k = 0; k < 2; k++) {
          double nk = g.fnormals[face * 2 + k];
          if (op == GRMP_OP_RECON_ID_RT0)
            rc[3 * k + node][f] = 0.5 * fv * nk;
          else {
            rc[3 * k + node][2 * f] = 0.5 * fv * nk;
            double c12 = (n == 0) ? (-1.0 / 12) : (1.0 / 12);
            rc[3 * k + node][2 * f + 1] = c12 * fv * nk * (double)g.signs[cell * 3 + f];
          }
        }
      }
      if (op == GRMP_OP_RECON_ID_RT0) rc[6 + f][f] = fv; else rc[6 + f][2 * f] = fv;
    }
  } else {
    const double B[3][3] = {{-1.0 / 36, -1.0 / 36, 1.0 / 18}, {-1.0 / 36, 1.0 / 18, -1.0 / 36}, {1.0 / 18, -1.0 / 36, -1.0 / 36}};
    const int r1[4] = {2, 2, 3, 1}, r2[4] = {1, 3, 1, 2};
    for (int f = 0; f < 4; f++) {
      i64 face = cf[f] - 1;
      double fv = g.fvol[face];
      for (int k = 0; k < 3; k++) {
        double nk = g.fnormals[face * 3 + k];
        for (int n = 0; n < 3; n++) {
          int node = TET_FACE[f][n];
          if (op == GRMP_OP_RECON_ID_RT0)
            rc[4 * k + node][f] = (1.0 / 3) * fv * nk;
          else {
            int o = g.orient[cell * 4 + f] - 1;
            rc[4 * k + node][3 * f] = (1.0 / 3) * nk * fv;
            rc[4 * k + node][3 * f + 1] = B[n][r1[o] - 1] * nk * fv;
            rc[4 * k + node][3 * f + 2] = B[n][r2[o] - 1] * nk * fv;
          }
        }
      }
      if (op == GRMP_OP_RECON_ID_RT0) rc[12 + f][f] = fv; else rc[12 + f][3 * f] = fv;
    }
  }
}

__device__ __forceinline__ bool is_recon(int op) { return op == GRMP_OP_RECON_ID_RT0 || op == GRMP_OP_RECON_ID_BDM1; }

// cvals[:, :, i] of evaluator e on the current cell (one quadrature point), reference loop order
template <int NDMAX, bool RECON>
__device__ void eval_qp(const EvalView& e, int edim, const Geo& T, const CellCoef<NDMAX>& cc, const double (*rc)[12], int i,
                        double (*cv)[NDMAX]) {
  const int nd = e.nd, nc = e.ncomp;
  const double* rv = e.refvals + (size_t)i * e.tab_nd * e.tab_nc;              // [dof][comp]
  const double* rdv = e.refderivs ? e.refderivs + (size_t)i * edim * e.tab_nd * e.tab_nc : nullptr;  // [j][row]
  const int nrow = e.tab_nd * e.tab_nc;
  if (RECON && is_recon(e.op)) {  // feevaluator_h1.jl:342-381
    double te[3][12];
    for (int dof = 0; dof < e.nd2; dof++)
      for (int k = 0; k < nc; k++) {
        double acc = 0.0;
        for (int l = 0; l < nc; l++) acc += T.A[k][l] * rv[cc.subset[dof] * e.tab_nc + l];
        acc *= cc.co[k][dof] / T.det;
        te[k][dof] = acc;
      }
    for (int di = 0; di < nd; di++)
      for (int k = 0; k < nc; k++) {
        double acc = 0.0;
        for (int dj = 0; dj < e.nd2; dj++)
          if (rc[di][dj] != 0) acc += rc[di][dj] * te[k][dj];
        cv[k][di] = acc;
      }
    return;
  }
  if (e.op == GRMP_OP_NORMALFLUX) {  // feevaluator_hdiv.jl:42-50: refbasisvals / |F|
    for (int dof = 0; dof < nd; dof++) cv[0][dof] = rv[dof * e.tab_nc] / T.det;
    return;
  }
  if (e.fam == FAM_RT0 || e.fam == FAM_BDM1) {
    if (e.op == GRMP_OP_ID) {  // feevaluator_hdiv.jl:2-19
      for (int dof = 0; dof < nd; dof++)
        for (int k = 0; k < edim; k++) {
          double acc = 0.0;
          for (int l = 0; l < edim; l++) acc += T.A[k][l] * rv[cc.subset[dof] * e.tab_nc + l];
          acc *= cc.co[k][dof] / T.det;
          cv[k][dof] = acc;
        }
    } else {  // Divergence, feevaluator_hdiv.jl:54-71
      for (int dof = 0; dof < nd; dof++) {
        double acc = 0.0;
        for (int j = 0; j < edim; j++) acc += rdv[j * nrow + cc.subset[dof] + j * e.tab_nd];
        acc *= cc.co[0][dof] / T.det;
        cv[0][dof] = acc;
      }
    }
    return;
  }
  const bool coeffs = (e.fam == FAM_H1BR);
  switch (e.op) {
    case GRMP_OP_ID:  // feevaluator.jl:100-103 / feevaluator_h1.jl:33-43
      for (int dof = 0; dof < nd; dof++)
        for (int k = 0; k < nc; k++) cv[k][dof] = coeffs ? rv[dof * e.tab_nc + k] * cc.co[k][dof] : rv[dof * e.tab_nc + k];
      break;
    case GRMP_OP_GRAD:  // feevaluator_h1.jl:61-74 / 77-94
      for (int dof = 0; dof < nd; dof++)
        for (int c = 0; c < nc; c++)
          for (int k = 0; k < edim; k++) {
            double acc = 0.0;
            for (int j = 0; j < edim; j++) acc += T.Ainv[k][j] * rdv[j * nrow + dof + c * e.tab_nd];
            if (coeffs) acc *= cc.co[c][dof];
            cv[k + c * edim][dof] = acc;
          }
      break;
    case GRMP_OP_SYMGRAD: {  // feevaluator_h1.jl:97-116 (offdiagval = 1), Voigt targets feevaluator.jl:231
      const int c2[4] = {0, 2, 2, 1};
      const int c3[9] = {0, 5, 4, 5, 1, 3, 4, 3, 2};
      const int nv = (edim == 2) ? 3 : 6;
      for (int dof = 0; dof < nd; dof++) {
        for (int v = 0; v < nv; v++) cv[v][dof] = 0.0;
        for (int c = 0; c < nc; c++)
          for (int k = 0; k < edim; k++)
            for (int j = 0; j < edim; j++) {
              int tgt = (edim == 2) ? c2[k + c * edim] : c3[k + c * edim];
              cv[tgt][dof] += T.Ainv[k][j] * rdv[j * nrow + dof + c * e.tab_nd];
            }
      }
      break;
    }
    case GRMP_OP_DIV:  // feevaluator_h1.jl:119-130 / 133-145
      for (int dof = 0; dof < nd; dof++) {
        double acc = 0.0;
        for (int k = 0; k < edim; k++)
          for (int j = 0; j < edim; j++) {
            if (coeffs) acc += T.Ainv[k][j] * rdv[j * nrow + dof + k * e.tab_nd] * cc.co[k][dof];
            else acc += T.Ainv[k][j] * rdv[j * nrow + dof + k * e.tab_nd];
          }
        cv[0][dof] = acc;
      }
      break;
  }
}

__device__ __forceinline__ bool needs_inverse(const EvalView& e) {
  return e.op == GRMP_OP_GRAD || e.op == GRMP_OP_SYMGRAD || (e.op == GRMP_OP_DIV && (e.fam == FAM_H1 || e.fam == FAM_H1BR));
}

__device__ __forceinline__ void apply_action(int action, const double* p, const double* in, double* out) {
  if (action == GRMP_ACT_HOOKE2D) {  // pdeoperators.jl:265-270
    double mu = p[0], la = p[1];
    out[0] = (la + 2 * mu) * in[0] + la * in[1];
    out[1] = (la + 2 * mu) * in[1] + la * in[0];
    out[2] = mu * in[2];
  } else {  // pdeoperators.jl:304-312
    double mu = p[0], la = p[1];
    out[0] = (la + 2 * mu) * in[0] + la * (in[1] + in[2]);
    out[1] = (la + 2 * mu) * in[1] + la * (in[0] + in[2]);
    out[2] = (la + 2 * mu) * in[2] + la * (in[0] + in[1]);
    out[3] = mu * in[3];
    out[4] = mu * in[4];
    out[5] = mu * in[5];
  }
}

template <int NDMAX, bool RECON>
__global__ void __launch_bounds__(128) blf_local_kernel(const BlfLocalParams p) {
  const i64 cell = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (cell >= p.g.ncells) return;
  const int nd1 = p.e1.nd, nd2 = p.e2.nd, nloc = nd1 * nd2, edim = p.g.dim;
  const i64 ncells = p.g.ncells;
  if (!cell_active(p.g, p.reg, cell)) {
    if (p.keys)
      for (int e = 0; e < nloc; e++) p.keys[cell * nloc + e] = ~0ull;
    return;
  }
  Geo T;
  geo_update(p.g, cell, T);
  if (needs_inverse(p.e1) || needs_inverse(p.e2)) geo_mapderiv(p.g, cell, T);
  CellCoef<NDMAX> cc1, cc2;
  double rc1[RECON ? NDMAX : 1][12], rc2[RECON ? NDMAX : 1][12];
  {
    const bool r1 = is_recon(p.e1.op);
    cell_coefficients<NDMAX>(p.g, cell, r1 ? p.e1.rfam : p.e1.fam, r1 ? p.e1.nd2 : nd1, p.e1.ncomp, cc1);
    if (RECON && r1) {
      for (int a = 0; a < nd1; a++)
        for (int b = 0; b < 12; b++) rc1[a][b] = 0.0;
      recon_coefficients<NDMAX>(p.g, cell, p.e1.op, rc1);
    }
    if (!p.same_eval) {
      const bool r2 = is_recon(p.e2.op);
      cell_coefficients<NDMAX>(p.g, cell, r2 ? p.e2.rfam : p.e2.fam, r2 ? p.e2.nd2 : nd2, p.e2.ncomp, cc2);
      if (RECON && r2) {
        for (int a = 0; a < nd2; a++)
          for (int b = 0; b < 12; b++) rc2[a][b] = 0.0;
        recon_coefficients<NDMAX>(p.g, cell, p.e2.op, rc2);
      }
    }
  }
  double loc[NDMAX * NDMAX];
  for (int e = 0; e < nloc; e++) loc[e] = 0.0;
  double cv1[RDMAX][NDMAX], cv2s[RDMAX][NDMAX];
  double(*cv2)[NDMAX] = p.same_eval ? cv1 : cv2s;
  const int rdim = p.e2.rd;
  const bool sym = (p.apt == GRMP_APT_SYMMETRIC);
  for (int i = 0; i < p.nq; i++) {
    eval_qp<NDMAX, RECON>(p.e1, edim, T, cc1, rc1, i, cv1);
    if (!p.same_eval) eval_qp<NDMAX, RECON>(p.e2, edim, T, cc2, rc2, i, cv2s);
    const double wi = p.w[i];
    for (int di = 0; di < nd1; di++) {
      double ar[RDMAX];
      if (p.action == GRMP_ACT_NONE) {
        for (int k = 0; k < rdim; k++) ar[k] = cv1[k][di];
      } else if (p.action == GRMP_ACT_CONVECTION) {   // pdeoperators.jl:459-467 on [a(x_i), operator evaluation of dof di]
        const double* a = p.aq + ((size_t)cell * p.nq + i) * p.aq_rd;
        for (int j = 0; j < rdim; j++) {
          double r = 0.0;
          for (int k = 0; k < p.aq_rd; k++) r += a[k] * cv1[j * p.aq_rd + k][di];
          ar[j] = r;
        }
      } else {
        double in[RDMAX];
        for (int k = 0; k < p.e1.rd; k++) in[k] = cv1[k][di];
        apply_action(p.action, p.act_p, in, ar);
      }
      if (p.apt == GRMP_APT_LUMPED) {
        double t = 0.0;
        for (int k = 0; k < rdim; k++) t += ar[k] * cv2[k][di];
        loc[di * nd2 + di] += wi * t;
      } else {
        for (int dj = (sym ? di : 0); dj < nd2; dj++) {
          double t = 0.0;
          for (int k = 0; k < rdim; k++) t += ar[k] * cv2[k][dj];
          loc[di * nd2 + dj] += wi * t;
        }
      }
    }
  }
  const double itemfactor = p.g.vol[cell] * p.factor * 1.0;  // bilinearform.jl:320
  const i32* d1 = p.e1.celldofs + cell * nd1;
  const i32* d2 = p.e2.celldofs + cell * nd2;
  for (int di = 0; di < nd1; di++)
    for (int dj = 0; dj < nd2; dj++) {
      if (sym && dj < di) continue;
      const double v = loc[di * nd2 + dj] * itemfactor;
      if (p.keys) {
        u64 key = ~0ull, keym = ~0ull;
        if (v != 0) {
          i64 r = d1[di] - 1, c = d2[dj] - 1;
          if (!sym && p.transposed) { i64 t = r; r = c; c = t; }
          key = (u64)c * (u64)p.nrows_key + (u64)r;
          if (sym && dj != di) keym = (u64)(d2[di] - 1) * (u64)p.nrows_key + (u64)(d1[dj] - 1);
        }
        p.keys[cell * nloc + di * nd2 + dj] = key;
        if (sym && dj != di) p.keys[cell * nloc + dj * nd2 + di] = keym;
      } else {
        p.lbuf[(i64)(di * nd2 + dj) * ncells + cell] = v;
      }
    }
}

template <int NDMAX, bool RECON>
__global__ void __launch_bounds__(128) lf_local_kernel(const LfLocalParams p) {
  const i64 cell = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (cell >= p.g.ncells) return;
  const int nd = p.e.nd, edim = p.g.dim;
  const i64 ncells = p.g.ncells;
  const bool act = cell_active(p.g, p.reg, cell);
  p.active[cell] = act ? 1 : 0;
  if (!act) return;
  Geo T;
  geo_update(p.g, cell, T);
  if (needs_inverse(p.e)) geo_mapderiv(p.g, cell, T);
  CellCoef<NDMAX> cc;
  double rc[RECON ? NDMAX : 1][12];
  const bool r = is_recon(p.e.op);
  cell_coefficients<NDMAX>(p.g, cell, r ? p.e.rfam : p.e.fam, r ? p.e.nd2 : nd, p.e.ncomp, cc);
  if (RECON && r) {
    for (int a = 0; a < nd; a++)
      for (int b = 0; b < 12; b++) rc[a][b] = 0.0;
    recon_coefficients<NDMAX>(p.g, cell, p.e.op, rc);
  }
  double lb[NDMAX];
  for (int d = 0; d < nd; d++) lb[d] = 0.0;
  double cv[RDMAX][NDMAX];
  const int rdim = p.e.rd;
  for (int i = 0; i < p.nq; i++) {
    eval_qp<NDMAX, RECON>(p.e, edim, T, cc, rc, i, cv);
    double f[RDMAX];
    for (int k = 0; k < rdim; k++)
      f[k] = (p.fsrc == GRMP_F_NONE) ? 1.0 : (p.fsrc == GRMP_F_CONST) ? p.fdata[k] : p.fdata[((size_t)cell * p.nq + i) * rdim + k];
    const double wi = p.w[i];
    for (int d = 0; d < nd; d++) {  // linearform.jl:181-210
      double t = 0.0;
      for (int k = 0; k < rdim; k++) t += f[k] * cv[k][d];
      lb[d] += t * wi;
    }
  }
  const double itemfactor = p.factor * p.g.vol[cell] * 1.0;  // linearform.jl:215
  for (int d = 0; d < nd; d++) p.lbuf[(i64)d * ncells + cell] = lb[d] * itemfactor;
}

// ItemIntegrator evaluate! (itemintegrator.jl:222-296): one thread per item, reference operation order.
//   input_i[k] = sum_dof coeffs[dof] * cvals[k,dof,i] * 1 from 0 in dof order (eval_febe!, feevaluator.jl:445-452)
//   NONE: result = input;  L2NORM (99-108): sum_j (0 + input[j])^2;  L2ERROR (52-69): sum_j (data[j] - input[j]*factor)^2
//   b[j,item] += result[j] * w_i * |T|  (266 / 291)
template <int NDMAX, bool RECON>
__global__ void __launch_bounds__(128) ii_local_kernel(const IiLocalParams p) {
  const i64 cell = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (cell >= p.g.ncells) return;
  const int nd = p.e.nd, edim = p.g.dim, rdim = p.e.rd, ardim = p.ardim;
  const i64 ncells = p.g.ncells;
  if (!cell_active(p.g, p.reg, cell)) {
    for (int j = 0; j < ardim; j++) p.itemval[(i64)j * ncells + cell] = 0.0;
    if (p.qtable)
      for (int k = 0; k < p.nq * rdim; k++) p.qtable[(size_t)cell * p.nq * rdim + k] = 0.0;
    return;
  }
  Geo T;
  geo_update(p.g, cell, T);
  if (needs_inverse(p.e)) geo_mapderiv(p.g, cell, T);
  CellCoef<NDMAX> cc;
  double rc[RECON ? NDMAX : 1][12];
  const bool r = is_recon(p.e.op);
  cell_coefficients<NDMAX>(p.g, cell, r ? p.e.rfam : p.e.fam, r ? p.e.nd2 : nd, p.e.ncomp, cc);
  if (RECON && r) {
    for (int a = 0; a < nd; a++)
      for (int b = 0; b < 12; b++) rc[a][b] = 0.0;
    recon_coefficients<NDMAX>(p.g, cell, p.e.op, rc);
  }
  double c[NDMAX];
  for (int d = 0; d < nd; d++) c[d] = p.coeffs[p.e.celldofs[cell * nd + d] - 1] * 1.0;
  double acc[RDMAX], own[RDMAX];
  for (int j = 0; j < ardim; j++) { acc[j] = p.b ? p.b[cell * ardim + j] : 0.0; own[j] = 0.0; }
  double cv[RDMAX][NDMAX];
  const double vol = p.g.vol[cell];
  for (int i = 0; i < p.nq; i++) {
    eval_qp<NDMAX, RECON>(p.e, edim, T, cc, rc, i, cv);
    double in[RDMAX], res[RDMAX];
    for (int k = 0; k < rdim; k++) in[k] = 0.0;
    for (int d = 0; d < nd; d++)
      for (int k = 0; k < rdim; k++) in[k] += c[d] * cv[k][d] * 1.0;
    if (p.qtable)
      for (int k = 0; k < rdim; k++) p.qtable[((size_t)cell * p.nq + i) * rdim + k] = in[k];
    if (p.kind == GRMP_II_NONE) {
      for (int j = 0; j < rdim; j++) res[j] = in[j];
    } else if (p.kind == GRMP_II_L2NORM) {
      double rr = 0.0;
      for (int j = 0; j < rdim; j++) { double t = 0.0; t += in[j]; rr += t * t; }
      res[0] = rr;
    } else {
      const double* dv = p.data + ((size_t)cell * p.nq + i) * rdim;
      double rr = 0.0;
      for (int j = 0; j < rdim; j++) { double v = dv[j]; v -= in[j] * p.factor; rr += v * v; }
      res[0] = rr;
    }
    const double wi = p.w[i];
    for (int j = 0; j < ardim; j++) {
      const double term = res[j] * wi * vol;
      acc[j] += term;
      own[j] += term;
    }
  }
  for (int j = 0; j < ardim; j++) {
    if (p.b) p.b[cell * ardim + j] = acc[j];
    p.itemval[(i64)j * ncells + cell] = own[j];
  }
}

int launch_ii_local(const IiLocalParams& p, cudaStream_t s) {
  const bool recon = (p.e.op == GRMP_OP_RECON_ID_RT0 || p.e.op == GRMP_OP_RECON_ID_BDM1);
  if (p.g.ncells == 0) return GRMP_OK;
  if (p.e.rd > RDMAX) return fail(GRMP_EUNSUPPORTED, "operator result longer than 9");
  const unsigned grid = (unsigned)((p.g.ncells + 127) / 128);
  if (recon) {
    if (p.e.nd > 16) return fail(GRMP_EUNSUPPORTED, "reconstruction operators support at most 16 local dofs");
    ii_local_kernel<16, true><<<grid, 128, 0, s>>>(p);
  } else if (p.e.nd <= 16)
    ii_local_kernel<16, false><<<grid, 128, 0, s>>>(p);
  else if (p.e.nd <= 30)
    ii_local_kernel<30, false><<<grid, 128, 0, s>>>(p);
  else
    return fail(GRMP_EUNSUPPORTED, "more than 30 local dofs per cell");
  GRMP_CUDA(cudaGetLastError());
  return GRMP_OK;
}

// NonlinearForm full_assemble! for the Newton convection form (nonlinearform.jl:114-233, kernel / jacobian pdeoperators.jl:459-488):
// one thread per cell, the oracle's (= the reference's) operation order.  in = [a_operator(u), ansatz_operator(u)](x_i) comes from the
// tables aq / gq; jac in2 runs over the stored columns of the sparse jacobian in ascending order.
template <int NDMAX>
__global__ void __launch_bounds__(128) nlf_local_kernel(const BlfLocalParams p) {
  const i64 cell = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (cell >= p.g.ncells) return;
  const int nd = p.e1.nd, nloc = nd * nd, edim = p.g.dim;
  const i64 ncells = p.g.ncells;
  const bool act = cell_active(p.g, p.reg, cell);
  if (p.active) p.active[cell] = act ? 1 : 0;
  if (!act) {
    if (p.keys)
      for (int e = 0; e < nloc; e++) p.keys[cell * nloc + e] = ~0ull;
    return;
  }
  Geo T;
  geo_update(p.g, cell, T);
  geo_mapderiv(p.g, cell, T);
  CellCoef<NDMAX> cc;
  double rcd[1][12];
  cell_coefficients<NDMAX>(p.g, cell, p.e1.fam, nd, p.e1.ncomp, cc);
  double loc[NDMAX * NDMAX], lb[NDMAX];
  for (int e = 0; e < nloc; e++) loc[e] = 0.0;
  for (int d = 0; d < nd; d++) lb[d] = 0.0;
  double cvg[RDMAX][NDMAX], cva[RDMAX][NDMAX], cvts[RDMAX][NDMAX];
  const int xdim = p.aq_rd, gdim = p.e1.rd, nc = p.e2.rd;
  const bool t_is_a = p.e2.op == p.ea.op, t_is_g = p.e2.op == p.e1.op;
  double(*cvt)[NDMAX] = t_is_a ? cva : (t_is_g ? cvg : cvts);
  for (int i = 0; i < p.nq; i++) {
    eval_qp<NDMAX, false>(p.e1, edim, T, cc, rcd, i, cvg);
    eval_qp<NDMAX, false>(p.ea, edim, T, cc, rcd, i, cva);
    if (!t_is_a && !t_is_g) eval_qp<NDMAX, false>(p.e2, edim, T, cc, rcd, i, cvts);
    const double* a = p.aq + ((size_t)cell * p.nq + i) * xdim;
    const double* gr = p.gq + ((size_t)cell * p.nq + i) * gdim;
    const double wi = p.w[i];
    double value[RDMAX], res[RDMAX];
    for (int j = 0; j < nc; j++) { double r = 0.0; for (int k = 0; k < xdim; k++) r += a[k] * gr[j * xdim + k]; value[j] = r; }
    for (int di = 0; di < nd; di++) {
      for (int j = 0; j < nc; j++) res[j] = 0.0;
      for (int k = 0; k < xdim; k++)
        for (int j = 0; j < nc; j++) res[j] += gr[j * xdim + k] * cva[k][di];
      for (int j = 0; j < nc; j++)
        for (int k = 0; k < xdim; k++) res[j] += a[k] * cvg[j * xdim + k][di];
      for (int dj = 0; dj < nd; dj++) {
        double t = 0.0;
        for (int k = 0; k < nc; k++) t += res[k] * cvt[k][dj];
        loc[di * nd + dj] += t * wi;
      }
    }
    for (int j = 0; j < nc; j++) res[j] = 0.0;
    for (int k = 0; k < xdim; k++)
      for (int j = 0; j < nc; j++) res[j] += gr[j * xdim + k] * a[k];
    for (int j = 0; j < nc; j++)
      for (int k = 0; k < xdim; k++) res[j] += a[k] * gr[j * xdim + k];
    for (int dj = 0; dj < nd; dj++) {
      double t = 0.0;
      for (int k = 0; k < nc; k++) t += (res[k] - value[k]) * cvt[k][dj];
      lb[dj] += t * wi;
    }
  }
  const double itemfactor = p.g.vol[cell] * p.factor * 1.0;
  const i32* d1 = p.e1.celldofs + cell * nd;
  for (int di = 0; di < nd; di++)
    for (int dj = 0; dj < nd; dj++) {
      const double l = loc[di * nd + dj];
      if (p.keys) {
        u64 key = ~0ull;
        if (l != 0) {       // _addnz(A, acol, arow, local, itemfactor): the test is on the unscaled entry
          i64 r = d1[di] - 1, c = d1[dj] - 1;
          if (p.transposed) { i64 t = r; r = c; c = t; }
          key = (u64)c * (u64)p.nrows_key + (u64)r;
        }
        p.keys[cell * nloc + di * nd + dj] = key;
      } else {
        p.lbuf[(i64)(di * nd + dj) * ncells + cell] = l * itemfactor;
      }
    }
  if (p.rbuf)
    for (int dj = 0; dj < nd; dj++) p.rbuf[(i64)dj * ncells + cell] = lb[dj] * itemfactor;
}

int launch_blf_local(const BlfLocalParams& p, cudaStream_t s) {
  const int nmax = p.e1.nd > p.e2.nd ? p.e1.nd : p.e2.nd;
  const bool recon = (p.e1.op == GRMP_OP_RECON_ID_RT0 || p.e1.op == GRMP_OP_RECON_ID_BDM1 || p.e2.op == GRMP_OP_RECON_ID_RT0 ||
                      p.e2.op == GRMP_OP_RECON_ID_BDM1);
  if (p.g.ncells == 0) return GRMP_OK;
  const unsigned grid = (unsigned)((p.g.ncells + 127) / 128);
  if (p.action == GRMP_ACT_NEWTON_CONVECTION) {
    if (recon || p.e1.fam == FAM_RT0 || p.e1.fam == FAM_BDM1) return fail(GRMP_EUNSUPPORTED, "Newton convection form: H1 spaces only");
    if (nmax <= 16) nlf_local_kernel<16><<<grid, 128, 0, s>>>(p);
    else if (nmax <= 30) nlf_local_kernel<30><<<grid, 128, 0, s>>>(p);
    else return fail(GRMP_EUNSUPPORTED, "more than 30 local dofs per cell");
    GRMP_CUDA(cudaGetLastError());
    return GRMP_OK;
  }
  if (recon) {
    if (nmax > 16) return fail(GRMP_EUNSUPPORTED, "reconstruction operators support at most 16 local dofs");
    blf_local_kernel<16, true><<<grid, 128, 0, s>>>(p);
  } else if (nmax <= 6)
    blf_local_kernel<6, false><<<grid, 128, 0, s>>>(p);
  else if (nmax <= 10)
    blf_local_kernel<10, false><<<grid, 128, 0, s>>>(p);
  else if (nmax <= 16)
    blf_local_kernel<16, false><<<grid, 128, 0, s>>>(p);
  else if (nmax <= 30)
    blf_local_kernel<30, false><<<grid, 128, 0, s>>>(p);
  else
    return fail(GRMP_EUNSUPPORTED, "more than 30 local dofs per cell");
  GRMP_CUDA(cudaGetLastError());
  return GRMP_OK;
}

int launch_lf_local(const LfLocalParams& p, cudaStream_t s) {
  const bool recon = (p.e.op == GRMP_OP_RECON_ID_RT0 || p.e.op == GRMP_OP_RECON_ID_BDM1);
  if (p.g.ncells == 0) return GRMP_OK;
  const unsigned grid = (unsigned)((p.g.ncells + 127) / 128);
  if (recon) {
    if (p.e.nd > 16) return fail(GRMP_EUNSUPPORTED, "reconstruction operators support at most 16 local dofs");
    lf_local_kernel<16, true><<<grid, 128, 0, s>>>(p);
  } else if (p.e.nd <= 16)
    lf_local_kernel<16, false><<<grid, 128, 0, s>>>(p);
  else if (p.e.nd <= 30)
    lf_local_kernel<30, false><<<grid, 128, 0, s>>>(p);
  else
    return fail(GRMP_EUNSUPPORTED, "more than 30 local dofs per cell");
  GRMP_CUDA(cudaGetLastError());
  return GRMP_OK;
}

}  // namespace grmp
