// colpath_ev.cuh -- per-cell affine pullbacks of the evaluators, written for the owner-computes column kernels
// (colpath.cu) and the cell-parallel scatter kernels (cellpath.cu).
//
// A local matrix entry of assemble! (bilinearform.jl:294-317) is
//     local[r, c] = sum_q w_q  cvR[:, r, q] . C . cvC[:, c, q]          (C: identity or the Hooke tensor)
// and every evaluator on the path is an affine image of a reference table,
//     cv[k, l, q] = sum_a J[k][a](cell) * coef[l](cell) * T[l][a][q]
// (feevaluator_h1.jl:61-145: J = L2GAinv; feevaluator_hdiv.jl:2-19, 54-71: J = A / det, coef = +-1).
// A thread that owns a matrix column evaluates ITS function once per quadrature point (table index lane-varying, from shared
// memory), applies weight, item factor and action, pulls the result back through the row evaluator's J,
//     U[a] = sum_k J[k][a] Y[k],
// and then every row is a short dot product of U with the ROW table, whose index is uniform over the warp (constant memory):
//     local[r, c] = sum_q sum_a U_q[a] * T_R[r][a][q].
// Vector-valued H1 spaces are componentwise copies of a scalar space (h1_p1.jl:64-75, h1_p2.jl:208-239) plus, for
// Bernardi-Raugel, one bubble per face multiplied by the face normal (h1v_br.jl:117-162, 218-273): tables hold the scalar
// functions only, beta[c] carries the component structure.
#pragma once
#include "common.cuh"

namespace grmp {

constexpr int CT_PAD = 16;          // column tables: 16 functions per (a, q) row = 128 B -> conflict-free for any lane pattern
constexpr int TABR_MAX = 4096;      // doubles of the row table in constant memory
constexpr int WQ_MAX = 64;

template <int ED> struct CellGeo {
  double A[ED][ED], Ainv[ED][ED], det, idet;
};

// update_trafo! / mapderiv! (feevaluator.jl:371-390): A[:,j] = x_{j+1} - x_1, Ainv = A^{-T}
template <int ED> __device__ __forceinline__ void cell_geo(const GridView& g, i64 cell, CellGeo<ED>& T) {
  i32 cn[ED + 1];
  if constexpr (ED == 3) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(g.cellnodes) + cell);
    cn[0] = v.x; cn[1] = v.y; cn[2] = v.z; cn[3] = v.w;
  } else {
#pragma unroll
    for (int j = 0; j < ED + 1; j++) cn[j] = __ldg(g.cellnodes + cell * (ED + 1) + j);
  }
  const double* x0 = g.coords + (i64)(cn[0] - 1) * ED;
  double b[ED];
#pragma unroll
  for (int k = 0; k < ED; k++) b[k] = x0[k];
#pragma unroll
  for (int j = 0; j < ED; j++) {
    const double* xj = g.coords + (i64)(cn[j + 1] - 1) * ED;
#pragma unroll
    for (int k = 0; k < ED; k++) T.A[k][j] = xj[k] - b[k];
  }
  const double(*A)[ED] = T.A;
  if constexpr (ED == 2) {
    T.det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
    T.idet = 1.0 / T.det;
    T.Ainv[1][1] = A[0][0] * T.idet;
    T.Ainv[1][0] = -A[0][1] * T.idet;
    T.Ainv[0][1] = -A[1][0] * T.idet;
    T.Ainv[0][0] = A[1][1] * T.idet;
  } else {
    const double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1], c01 = A[1][0] * A[2][2] - A[1][2] * A[2][0], c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
    T.det = A[0][0] * c00 - A[0][1] * c01 + A[0][2] * c02;
    T.idet = 1.0 / T.det;
    T.Ainv[0][0] = c00 * T.idet;
    T.Ainv[0][1] = -c01 * T.idet;
    T.Ainv[0][2] = c02 * T.idet;
    T.Ainv[1][0] = -(A[0][1] * A[2][2] - A[0][2] * A[2][1]) * T.idet;
    T.Ainv[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * T.idet;
    T.Ainv[1][2] = -(A[0][0] * A[2][1] - A[0][1] * A[2][0]) * T.idet;
    T.Ainv[2][0] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * T.idet;
    T.Ainv[2][1] = -(A[0][0] * A[1][2] - A[0][2] * A[1][0]) * T.idet;
    T.Ainv[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * T.idet;
  }
}

__host__ __device__ constexpr int voigt_slot(int ed, int kc) {   // feevaluator.jl:231: [1,3,3,2] / [1,6,5,6,2,4,5,4,3], 0-based here
  return ed == 2 ? (kc == 0 ? 0 : (kc == 3 ? 1 : 2))
                 : (kc == 0 ? 0 : kc == 1 ? 5 : kc == 2 ? 4 : kc == 3 ? 5 : kc == 4 ? 1 : kc == 5 ? 3 : kc == 6 ? 4 : kc == 7 ? 3 : 2);
}

// ---- componentwise H1 spaces: P1 / P2 / P0 with NC components, Bernardi-Raugel (NC = ED, NBUB = ED + 1) -------------------
template <int ED_, int NC_, int NDS_, int NBUB_, int OP_> struct H1Ev {
  static constexpr int ED = ED_, NC = NC_, NDS = NDS_, NBUB = NBUB_, OP = OP_, KIND = 0;
  static constexpr int ND = NC * NDS + NBUB, NROW = ND, NSF = NDS + NBUB;
  static constexpr bool DER = (OP != GRMP_OP_ID);
  static constexpr int NAS = DER ? ED : 1, NCU = NC;
  static constexpr int RD = OP == GRMP_OP_ID ? NC : OP == GRMP_OP_GRAD ? NC * ED : OP == GRMP_OP_SYMGRAD ? (ED == 2 ? 3 : 6) : 1;
  static constexpr int NM = DER ? ED * ED : 0;
  static constexpr int CACHE_N = NM + NBUB * ED;
  static_assert(NSF <= CT_PAD, "column table row too short");
  struct Regs {
    double M[NM > 0 ? NM : 1];
    double nb[NBUB > 0 ? NBUB * ED : 1];
  };
  __device__ __forceinline__ static void build_cache(const GridView& g, i64 cell, const CellGeo<ED>& T, double* out) {
    if constexpr (DER) {
#pragma unroll
      for (int k = 0; k < ED; k++)
#pragma unroll
        for (int a = 0; a < ED; a++) out[k * ED + a] = T.Ainv[k][a];
    }
    if constexpr (NBUB > 0) {   // h1v_br.jl:150-162, 253-273: bubble coefficients = face normal
      const i32* cf = g.cellfaces + cell * (ED + 1);
#pragma unroll
      for (int b = 0; b < NBUB; b++)
#pragma unroll
        for (int c = 0; c < ED; c++) out[NM + b * ED + c] = g.fnormals[(i64)(cf[b] - 1) * ED + c];
    }
  }
  __device__ __forceinline__ static void load(const double* cr, Regs& R) {
#pragma unroll
    for (int i = 0; i < NM; i++) R.M[i] = cr[i];
#pragma unroll
    for (int i = 0; i < NBUB * ED; i++) R.nb[i] = cr[NM + i];
  }
  // what does not depend on the quadrature point: scalar function and component weights of local function l
  struct Prep {
    int s;
    double beta[NC];
  };
  __device__ __forceinline__ static void prep(const Regs& R, int l, Prep& P) {
    if (NBUB > 0 && l >= NC * NDS) {
      const int b = l - NC * NDS;
      P.s = NDS + b;
#pragma unroll
      for (int c = 0; c < NC; c++) {
        double v = 0.0;
#pragma unroll
        for (int bb = 0; bb < NBUB; bb++) v = (bb == b) ? R.nb[bb * ED + c] : v;
        P.beta[c] = v;
      }
    } else {
      const int cl = l / NDS;
      P.s = l - cl * NDS;
#pragma unroll
      for (int c = 0; c < NC; c++) P.beta[c] = (c == cl) ? 1.0 : 0.0;
    }
  }
  // operator values of local function l at quadrature point q (Ct: [a][q][CT_PAD] scalar table)
  __device__ __forceinline__ static void col_eval(const Regs& R, const Prep& P, const double* __restrict__ Ct, int nq, int q, double (&Y)[RD]) {
    const int s = P.s;
    const double(&beta)[NC] = P.beta;
    if constexpr (OP == GRMP_OP_ID) {
      const double v = Ct[(size_t)q * CT_PAD + s];
#pragma unroll
      for (int c = 0; c < NC; c++) Y[c] = beta[c] * v;
    } else {
      double d[ED], gr[ED];
#pragma unroll
      for (int a = 0; a < ED; a++) d[a] = Ct[((size_t)a * nq + q) * CT_PAD + s];
#pragma unroll
      for (int k = 0; k < ED; k++) {
        double t = 0.0;
#pragma unroll
        for (int a = 0; a < ED; a++) t = fma(R.M[k * ED + a], d[a], t);
        gr[k] = t;
      }
      if constexpr (OP == GRMP_OP_GRAD) {
#pragma unroll
        for (int c = 0; c < NC; c++)
#pragma unroll
          for (int k = 0; k < ED; k++) Y[c * ED + k] = beta[c] * gr[k];
      } else if constexpr (OP == GRMP_OP_SYMGRAD) {   // feevaluator_h1.jl:97-116, offdiagval = 1
#pragma unroll
        for (int v = 0; v < RD; v++) Y[v] = 0.0;
#pragma unroll
        for (int c = 0; c < NC; c++)
#pragma unroll
          for (int k = 0; k < ED; k++) Y[voigt_slot(ED, k + c * ED)] += beta[c] * gr[k];
      } else {   // Divergence, feevaluator_h1.jl:119-145
        double t = 0.0;
#pragma unroll
        for (int c = 0; c < NC; c++) t = fma(beta[c], gr[c], t);
        Y[0] = t;
      }
    }
  }
  // U[c][a] = sum_k J[k][a] Y[...]: the column value pulled back to the reference directions of component c
  __device__ __forceinline__ static void pullback(const Regs& R, const double (&Y)[RD], double (&U)[NCU][NAS]) {
    if constexpr (OP == GRMP_OP_ID) {
#pragma unroll
      for (int c = 0; c < NC; c++) U[c][0] = Y[c];
    } else {
#pragma unroll
      for (int c = 0; c < NC; c++)
#pragma unroll
        for (int a = 0; a < ED; a++) {
          double t = 0.0;
          if constexpr (OP == GRMP_OP_GRAD) {
#pragma unroll
            for (int k = 0; k < ED; k++) t = fma(R.M[k * ED + a], Y[c * ED + k], t);
          } else if constexpr (OP == GRMP_OP_SYMGRAD) {
#pragma unroll
            for (int k = 0; k < ED; k++) t = fma(R.M[k * ED + a], Y[voigt_slot(ED, k + c * ED)], t);
          } else {
            t = R.M[c * ED + a] * Y[0];
          }
          U[c][a] = t;
        }
    }
  }
  // accumulators of one column: E[c][s] for the componentwise rows, Eb[b][c] for the bubble rows
  struct Acc {
    double E[NC][NDS];
    double Eb[NBUB > 0 ? NBUB : 1][NC];
  };
  __device__ __forceinline__ static void acc_zero(Acc& A) {
#pragma unroll
    for (int c = 0; c < NC; c++)
#pragma unroll
      for (int s = 0; s < NDS; s++) A.E[c][s] = 0.0;
#pragma unroll
    for (int b = 0; b < NBUB; b++)
#pragma unroll
      for (int c = 0; c < NC; c++) A.Eb[b][c] = 0.0;
  }
  // Rt: row table of quadrature point q, [s][a] (uniform address -> constant memory operand)
  __device__ __forceinline__ static void acc_rows(Acc& A, const double (&U)[NCU][NAS], const double* __restrict__ Rt) {
#pragma unroll
    for (int s = 0; s < NDS; s++)
#pragma unroll
      for (int a = 0; a < NAS; a++) {
        const double t = Rt[s * NAS + a];
#pragma unroll
        for (int c = 0; c < NC; c++) A.E[c][s] = fma(U[c][a], t, A.E[c][s]);
      }
#pragma unroll
    for (int b = 0; b < NBUB; b++)
#pragma unroll
      for (int a = 0; a < NAS; a++) {
        const double t = Rt[(NDS + b) * NAS + a];
#pragma unroll
        for (int c = 0; c < NC; c++) A.Eb[b][c] = fma(U[c][a], t, A.Eb[b][c]);
      }
  }
  template <class F> __device__ __forceinline__ static void emit_rows(const Regs& R, const Acc& A, F&& f) {
#pragma unroll
    for (int c = 0; c < NC; c++)
#pragma unroll
      for (int s = 0; s < NDS; s++) f(c * NDS + s, A.E[c][s]);
#pragma unroll
    for (int b = 0; b < NBUB; b++) {
      double t = 0.0;
#pragma unroll
      for (int c = 0; c < NC; c++) t = fma(R.nb[b * ED + c], A.Eb[b][c], t);
      f(NC * NDS + b, t);
    }
  }
};

// ---- Hdiv spaces (contravariant Piola): RT0 (NDALL = ED + 1), BDM1 (2D: 6, 3D: 16 reference functions of which 12 are selected) ----
template <int ED_, int NDALL_, int OP_> struct HdivEv {
  static constexpr int ED = ED_, NC = 1, NDS = NDALL_, NBUB = 0, OP = OP_, KIND = 1;
  static constexpr int NROW = NDALL_, NSF = NDALL_;
  static constexpr int NAS = OP == GRMP_OP_ID ? ED : 1, NCU = 1;
  static constexpr int RD = OP == GRMP_OP_ID ? ED : 1;
  static constexpr int NM = ED * ED;
  static constexpr int CACHE_N = NM + 2;
  static_assert(NSF <= CT_PAD, "column table row too short");
  struct Regs {
    double M[NM];
    double idet;
    u32 neg;
  };
  // bit r of the mask: coefficient of reference function r is negative (hdiv_rt0.jl:106-116, hdiv_bdm1.jl:278-307)
  __device__ __forceinline__ static u32 negmask(const GridView& g, i64 cell) {
    const i32* sg = g.signs + cell * (ED + 1);
    u32 m = 0;
    if constexpr (NDALL_ == ED + 1) {
#pragma unroll
      for (int j = 0; j < ED + 1; j++) m |= (sg[j] < 0 ? 1u : 0u) << j;
    } else if constexpr (ED == 2) {
#pragma unroll
      for (int j = 0; j < 3; j++) m |= (sg[j] < 0 ? 1u : 0u) << (2 * j);
    } else {   // 3D: local dof 3j -> ref 4j (sign), 3j+1 -> ref 4j+3-s1[o] (-1), 3j+2 -> ref 4j+3-s2[o] (+1)
      const i32* o = g.orient + cell * 4;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int oj = o[j] - 1;
        const int s1 = oj == 0 ? 1 : (oj == 1 ? 0 : (oj == 2 ? 1 : 2));
        m |= (sg[j] < 0 ? 1u : 0u) << (4 * j);
        m |= 1u << (4 * j + 3 - s1);
      }
    }
    return m;
  }
  __device__ __forceinline__ static void build_cache(const GridView& g, i64 cell, const CellGeo<ED>& T, double* out) {
#pragma unroll
    for (int k = 0; k < ED; k++)
#pragma unroll
      for (int a = 0; a < ED; a++) out[k * ED + a] = T.A[k][a];
    out[NM] = T.idet;
    out[NM + 1] = __longlong_as_double((long long)negmask(g, cell));
  }
  __device__ __forceinline__ static void load(const double* cr, Regs& R) {
#pragma unroll
    for (int i = 0; i < NM; i++) R.M[i] = cr[i];
    R.idet = cr[NM];
    R.neg = (u32)__double_as_longlong(cr[NM + 1]);
  }
  struct Prep {
    int l;
    double sg;
  };
  __device__ __forceinline__ static void prep(const Regs& R, int l, Prep& P) {
    P.l = l;
    P.sg = ((R.neg >> l) & 1u) ? -R.idet : R.idet;
  }
  __device__ __forceinline__ static void col_eval(const Regs& R, const Prep& P, const double* __restrict__ Ct, int nq, int q, double (&Y)[RD]) {
    const double sg = P.sg;
    const int l = P.l;
    if constexpr (OP == GRMP_OP_ID) {
      double d[ED];
#pragma unroll
      for (int a = 0; a < ED; a++) d[a] = Ct[((size_t)a * nq + q) * CT_PAD + l];
#pragma unroll
      for (int k = 0; k < ED; k++) {
        double t = 0.0;
#pragma unroll
        for (int a = 0; a < ED; a++) t = fma(R.M[k * ED + a], d[a], t);
        Y[k] = sg * t;
      }
    } else {
      Y[0] = sg * Ct[(size_t)q * CT_PAD + l];
    }
  }
  __device__ __forceinline__ static void pullback(const Regs& R, const double (&Y)[RD], double (&U)[NCU][NAS]) {
    if constexpr (OP == GRMP_OP_ID) {
#pragma unroll
      for (int a = 0; a < ED; a++) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < ED; k++) t = fma(R.M[k * ED + a], Y[k], t);
        U[0][a] = R.idet * t;
      }
    } else {
      U[0][0] = R.idet * Y[0];
    }
  }
  struct Acc {
    double E[NDALL_];
  };
  __device__ __forceinline__ static void acc_zero(Acc& A) {
#pragma unroll
    for (int r = 0; r < NDALL_; r++) A.E[r] = 0.0;
  }
  __device__ __forceinline__ static void acc_rows(Acc& A, const double (&U)[NCU][NAS], const double* __restrict__ Rt) {
#pragma unroll
    for (int r = 0; r < NDALL_; r++)
#pragma unroll
      for (int a = 0; a < NAS; a++) A.E[r] = fma(U[0][a], Rt[r * NAS + a], A.E[r]);
  }
  template <class F> __device__ __forceinline__ static void emit_rows(const Regs& R, const Acc& A, F&& f) {
#pragma unroll
    for (int r = 0; r < NDALL_; r++) f(r, ((R.neg >> r) & 1u) ? -A.E[r] : A.E[r]);
  }
};

// ---- ReconstructionIdentity{HDIVRT0{2} | HDIVBDM1{2}} on Bernardi-Raugel, triangles (feevaluator_h1.jl:342-381) ----------------
// R phi_l = sum_r rc[l][r] psi_r with psi_r the Piola-mapped Hdiv basis and rc from boundary_coefficients!
// (reconstructions.jl:353-403): for a nodal dof (component k, node n) and every face F containing n
//     rc[(k,n)][F, RT0 function] = 1/2 |F| n_k(F),      rc[(k,n)][F, BDM1 function] = -+1/12 |F| n_k(F) sign_F  (n first / second node of F)
// and rc[bubble F][F, RT0 function] = |F|.  The reconstruction is FUSED into the contraction: a column thread combines the
// reference functions with its weights before the Piola map, the row side accumulates against the Hdiv reference table
// (uniform index, constant memory) and applies rc when the rows are emitted -- the coefficient matrix is never formed.
template <int NDALL2_> struct ReconEv2D {
  static constexpr int ED = 2, NC = 1, NDS = NDALL2_, NBUB = 0, KIND = 2;
  static constexpr bool BDM = (NDALL2_ == 6);
  static constexpr int OP = BDM ? GRMP_OP_RECON_ID_BDM1 : GRMP_OP_RECON_ID_RT0;
  static constexpr int NF = 3, NN = 3, NM2 = BDM ? 2 : 1;
  static constexpr int NROW = ED * NN + NF, NSF = NDALL2_;
  static constexpr int NAS = ED, NCU = 1, RD = ED, NM = ED * ED;
  static constexpr int CACHE_N = NM + 2 + NF * ED + NF;
  struct Regs {
    double M[NM], idet, fw[NF * ED], fv[NF];
    u32 bits;     // bit f: CellFaceSigns[f] < 0
  };
  __device__ __forceinline__ static void build_cache(const GridView& g, i64 cell, const CellGeo<ED>& T, double* out) {
#pragma unroll
    for (int k = 0; k < ED; k++)
#pragma unroll
      for (int a = 0; a < ED; a++) out[k * ED + a] = T.A[k][a];
    out[NM] = T.idet;
    const i32* sg = g.signs + cell * NF;
    const i32* cf = g.cellfaces + cell * NF;
    u32 bits = 0;
#pragma unroll
    for (int f = 0; f < NF; f++) {
      bits |= (sg[f] < 0 ? 1u : 0u) << f;
      const i64 face = cf[f] - 1;
      const double fv = g.fvol[face];
#pragma unroll
      for (int k = 0; k < ED; k++) out[NM + 2 + f * ED + k] = fv * g.fnormals[face * ED + k];
      out[NM + 2 + NF * ED + f] = fv;
    }
    out[NM + 1] = __longlong_as_double((long long)bits);
  }
  __device__ __forceinline__ static void load(const double* cr, Regs& R) {
#pragma unroll
    for (int i = 0; i < NM; i++) R.M[i] = cr[i];
    R.idet = cr[NM];
    R.bits = (u32)__double_as_longlong(cr[NM + 1]);
#pragma unroll
    for (int i = 0; i < NF * ED; i++) R.fw[i] = cr[NM + 2 + i];
#pragma unroll
    for (int i = 0; i < NF; i++) R.fv[i] = cr[NM + 2 + NF * ED + i];
  }
  // local face -> nodes of Triangle2D: [1 2], [2 3], [3 1]
  __host__ __device__ static constexpr int face_node(int f, int pos) { return pos == 0 ? f : (f + 1) % 3; }
  // weight of Hdiv local dof (f, m) in R phi_l, including the Hdiv coefficient sign of that dof (hdiv_rt0.jl:106-116,
  // hdiv_bdm1.jl:278-290: CellFaceSigns on the RT0-type functions only)
  __device__ __forceinline__ static double weight(const Regs& R, int l, int f, int m) {
    const double sgn = ((R.bits >> f) & 1u) ? -1.0 : 1.0;
    double w = 0.0;
    if (l >= ED * NN) {
      if (l - ED * NN == f && m == 0) w = R.fv[f] * sgn;
    } else {
      const int k = l / NN, node = l - k * NN;
      const double fwk = k == 0 ? R.fw[f * ED] : R.fw[f * ED + 1];
      if (node == face_node(f, 0)) w = (m == 0) ? 0.5 * fwk * sgn : (-1.0 / 12.0) * fwk * sgn;    // BDM1 part: rc carries sign_F, the function none
      else if (node == face_node(f, 1)) w = (m == 0) ? 0.5 * fwk * sgn : (1.0 / 12.0) * fwk * sgn;
    }
    return w;
  }
  struct Prep {
    double w[NF][NM2];
  };
  __device__ __forceinline__ static void prep(const Regs& R, int l, Prep& P) {
#pragma unroll
    for (int f = 0; f < NF; f++)
#pragma unroll
      for (int m = 0; m < NM2; m++) P.w[f][m] = weight(R, l, f, m);
  }
  __device__ __forceinline__ static void col_eval(const Regs& R, const Prep& P, const double* __restrict__ Ct, int nq, int q, double (&Y)[RD]) {
    double h[ED] = {0.0, 0.0};
#pragma unroll
    for (int f = 0; f < NF; f++)
#pragma unroll
      for (int m = 0; m < NM2; m++) {
        const double w = P.w[f][m];
        const int r = BDM ? 2 * f + m : f;
#pragma unroll
        for (int a = 0; a < ED; a++) h[a] = fma(w, Ct[((size_t)a * nq + q) * CT_PAD + r], h[a]);
      }
#pragma unroll
    for (int k = 0; k < ED; k++) Y[k] = R.idet * (R.M[k * ED] * h[0] + R.M[k * ED + 1] * h[1]);
  }
  __device__ __forceinline__ static void pullback(const Regs& R, const double (&Y)[RD], double (&U)[NCU][NAS]) {
#pragma unroll
    for (int a = 0; a < ED; a++) U[0][a] = R.idet * (R.M[a] * Y[0] + R.M[ED + a] * Y[1]);
  }
  struct Acc {
    double E[NDALL2_];
  };
  __device__ __forceinline__ static void acc_zero(Acc& A) {
#pragma unroll
    for (int r = 0; r < NDALL2_; r++) A.E[r] = 0.0;
  }
  __device__ __forceinline__ static void acc_rows(Acc& A, const double (&U)[NCU][NAS], const double* __restrict__ Rt) {
#pragma unroll
    for (int r = 0; r < NDALL2_; r++)
#pragma unroll
      for (int a = 0; a < NAS; a++) A.E[r] = fma(U[0][a], Rt[r * NAS + a], A.E[r]);
  }
  template <class F> __device__ __forceinline__ static void emit_rows(const Regs& R, const Acc& A, F&& f) {
#pragma unroll
    for (int l = 0; l < NROW; l++) {
      double v = 0.0;
#pragma unroll
      for (int fc = 0; fc < NF; fc++)
#pragma unroll
        for (int m = 0; m < NM2; m++) v = fma(weight(R, l, fc, m), A.E[BDM ? 2 * fc + m : fc], v);
      f(l, v);
    }
  }
};

// Hooke tensors (pdeoperators.jl:265-270, 304-312) applied to the column value
template <int ACT, int RD> __device__ __forceinline__ void apply_action_col(const double* p, double (&Y)[RD]) {
  if constexpr (ACT == GRMP_ACT_HOOKE2D) {
    static_assert(RD == 3, "Hooke 2D acts on Voigt vectors of length 3");
    const double mu = p[0], la = p[1];
    const double a = (la + 2 * mu) * Y[0] + la * Y[1], b = (la + 2 * mu) * Y[1] + la * Y[0];
    Y[0] = a; Y[1] = b; Y[2] = mu * Y[2];
  } else if constexpr (ACT == GRMP_ACT_HOOKE3D) {
    static_assert(RD == 6, "Hooke 3D acts on Voigt vectors of length 6");
    const double mu = p[0], la = p[1];
    const double a = (la + 2 * mu) * Y[0] + la * (Y[1] + Y[2]), b = (la + 2 * mu) * Y[1] + la * (Y[0] + Y[2]),
                 c = (la + 2 * mu) * Y[2] + la * (Y[0] + Y[1]);
    Y[0] = a; Y[1] = b; Y[2] = c; Y[3] = mu * Y[3]; Y[4] = mu * Y[4]; Y[5] = mu * Y[5];
  }
}

// Geometry records (N doubles per cell, N even).  The lanes of a warp read records of different cells, so every load instruction costs one
// wavefront per lane on the L1 data pipe -- which is what bounds the column kernels.  Records of 32-byte-multiple size are 32-byte aligned
// and read with 256-bit loads (LDG.E.256 on sm_100a).  Records of N = 2 (mod 4) >= 10 doubles are stored split: the first N - 2 doubles of
// every cell as an array of aligned quads, the last pair in a second array behind it (one 16-byte load): ceil(N/4) load instructions
// instead of N/2, no padding bytes (P1 tetrahedron Laplacian, 24 pairs per column: 0.62 -> 0.57 ms).  Measured alternatives: padding
// every record to a multiple of 32 bytes loses (Bernardi-Raugel Laplacian -5 %, Hooke -4 %: more bytes than saved wavefronts), and
// splitting 6-double records loses 1.4 % on the Hooke form (a second cache line per pair for one saved instruction), so those stay
// contiguous with 16-byte loads.
template <int N> struct rec_layout {
  static constexpr bool SPLIT = (N % 4 == 2) && N >= 10;
  static constexpr int Q4 = SPLIT ? N - 2 : N;          // doubles of a cell in the first array
};
template <int N> __device__ __forceinline__ void load_record(const double* __restrict__ geo, i64 ncells, i64 cell, double (&cr)[N]) {
  static_assert(N % 2 == 0, "record stride");
  constexpr int Q4 = rec_layout<N>::Q4;
  const double* __restrict__ src = geo + cell * Q4;
  if constexpr (Q4 % 4 == 0) {
#pragma unroll
    for (int i = 0; i < Q4 / 4; i++)
      asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(cr[4 * i]), "=d"(cr[4 * i + 1]), "=d"(cr[4 * i + 2]), "=d"(cr[4 * i + 3]) : "l"(src + 4 * i));
  } else {
#pragma unroll
    for (int i = 0; i < Q4 / 2; i++) { const double2 t = __ldg(reinterpret_cast<const double2*>(src) + i); cr[2 * i] = t.x; cr[2 * i + 1] = t.y; }
  }
  if constexpr (rec_layout<N>::SPLIT) {
    const double2 t = __ldg(reinterpret_cast<const double2*>(geo + ncells * Q4) + cell);
    cr[Q4] = t.x; cr[Q4 + 1] = t.y;
  }
}
template <int N> __device__ __forceinline__ void store_record(double* __restrict__ geo, i64 ncells, i64 cell, const double (&cr)[N]) {
  constexpr int Q4 = rec_layout<N>::Q4;
  double2* dst = reinterpret_cast<double2*>(geo + cell * Q4);
#pragma unroll
  for (int i = 0; i < Q4 / 2; i++) dst[i] = make_double2(cr[2 * i], cr[2 * i + 1]);
  if constexpr (rec_layout<N>::SPLIT) reinterpret_cast<double2*>(geo + ncells * Q4)[cell] = make_double2(cr[Q4], cr[Q4 + 1]);
}

// cache record of one cell: [0] item factor (CellVolumes * factor, bilinearform.jl:320), then the row evaluator's part, then
// the column evaluator's part unless both need the same data
template <class RowEv, class ColEv> struct CacheLayout {
  static constexpr bool SAME = (RowEv::KIND == ColEv::KIND && RowEv::NC == ColEv::NC && RowEv::NDS == ColEv::NDS && RowEv::NBUB == ColEv::NBUB &&
                                RowEv::CACHE_N == ColEv::CACHE_N);
  static constexpr int OFF_R = 1, OFF_C = SAME ? 1 : 1 + RowEv::CACHE_N;
  static constexpr int N = 1 + RowEv::CACHE_N + (SAME ? 0 : ColEv::CACHE_N);
  static constexpr int STRIDE = (N + 1) & ~1;     // even: records stay 16-byte aligned (32-byte aligned when a multiple of 4: load_record)
};

// (the column kernels store CellVolumes here and apply `factor` when they read the record)
template <class RowEv, class ColEv>
__device__ __forceinline__ void build_cell_cache(const GridView& g, i64 cell, double factor, double* cr) {
  using L = CacheLayout<RowEv, ColEv>;
  CellGeo<RowEv::ED> T;
  cell_geo<RowEv::ED>(g, cell, T);
  cr[0] = g.vol[cell] * factor;
  RowEv::build_cache(g, cell, T, cr + L::OFF_R);
  if constexpr (!L::SAME) ColEv::build_cache(g, cell, T, cr + L::OFF_C);
}

// forms with a kernel: X(row evaluator, column evaluator, action, NQ).  NQ is the number of quadrature points of the rule that
// prepare_assembly! picks for the form (assemblypatterns.jl:559-565: order = sum(polynomial order + operator shift); midpoint
// rules 1 point, order-2 rules 3 / 4 points, order 4: 9-point Stroud rule / 11-point tetrahedron rule).  A compile-time NQ lets
// the row table sit in the FMA's constant operand; forms assembled with another rule (bonus_quadorder) use the generic path.
#define GRMP_H1_SQUARE(X, ED, NC, NDS, NB, NQG, NQI)                                                             \
  X((H1Ev<ED, NC, NDS, NB, GRMP_OP_GRAD>), (H1Ev<ED, NC, NDS, NB, GRMP_OP_GRAD>), GRMP_ACT_NONE, NQG)            \
  X((H1Ev<ED, NC, NDS, NB, GRMP_OP_ID>), (H1Ev<ED, NC, NDS, NB, GRMP_OP_ID>), GRMP_ACT_NONE, NQI)
#define GRMP_HDIV_SQUARE(X, ED, NDALL, NQI)                                                                      \
  X((HdivEv<ED, NDALL, GRMP_OP_ID>), (HdivEv<ED, NDALL, GRMP_OP_ID>), GRMP_ACT_NONE, NQI)                        \
  X((HdivEv<ED, NDALL, GRMP_OP_DIV>), (HdivEv<ED, NDALL, GRMP_OP_DIV>), GRMP_ACT_NONE, 1)
#define GRMP_RECT(X, A, B, NQ) X(A, B, GRMP_ACT_NONE, NQ) X(B, A, GRMP_ACT_NONE, NQ)

#define GRMP_SQUARE_FORMS(X)                                                                                     \
  GRMP_H1_SQUARE(X, 2, 1, 3, 0, 1, 3) GRMP_H1_SQUARE(X, 2, 2, 3, 0, 1, 3) GRMP_H1_SQUARE(X, 2, 1, 6, 0, 3, 9)    \
  GRMP_H1_SQUARE(X, 2, 2, 6, 0, 3, 9) GRMP_H1_SQUARE(X, 2, 2, 3, 3, 3, 9)                                        \
  X((H1Ev<2, 1, 1, 0, GRMP_OP_ID>), (H1Ev<2, 1, 1, 0, GRMP_OP_ID>), GRMP_ACT_NONE, 1)                            \
  GRMP_H1_SQUARE(X, 3, 1, 4, 0, 1, 4) GRMP_H1_SQUARE(X, 3, 3, 4, 0, 1, 4) GRMP_H1_SQUARE(X, 3, 1, 10, 0, 4, 11)  \
  GRMP_H1_SQUARE(X, 3, 3, 10, 0, 4, 11)                                                                          \
  X((H1Ev<3, 3, 4, 4, GRMP_OP_GRAD>), (H1Ev<3, 3, 4, 4, GRMP_OP_GRAD>), GRMP_ACT_NONE, 11)                       \
  X((H1Ev<3, 1, 1, 0, GRMP_OP_ID>), (H1Ev<3, 1, 1, 0, GRMP_OP_ID>), GRMP_ACT_NONE, 1)                            \
  X((H1Ev<2, 2, 3, 0, GRMP_OP_SYMGRAD>), (H1Ev<2, 2, 3, 0, GRMP_OP_SYMGRAD>), GRMP_ACT_HOOKE2D, 1)               \
  X((H1Ev<2, 2, 6, 0, GRMP_OP_SYMGRAD>), (H1Ev<2, 2, 6, 0, GRMP_OP_SYMGRAD>), GRMP_ACT_HOOKE2D, 3)               \
  X((H1Ev<3, 3, 4, 0, GRMP_OP_SYMGRAD>), (H1Ev<3, 3, 4, 0, GRMP_OP_SYMGRAD>), GRMP_ACT_HOOKE3D, 1)               \
  X((H1Ev<3, 3, 10, 0, GRMP_OP_SYMGRAD>), (H1Ev<3, 3, 10, 0, GRMP_OP_SYMGRAD>), GRMP_ACT_HOOKE3D, 4)             \
  GRMP_HDIV_SQUARE(X, 2, 3, 3) GRMP_HDIV_SQUARE(X, 2, 6, 3) GRMP_HDIV_SQUARE(X, 3, 4, 4) GRMP_HDIV_SQUARE(X, 3, 16, 4)  \
  X((ReconEv2D<3>), (ReconEv2D<3>), GRMP_ACT_NONE, 9) X((ReconEv2D<6>), (ReconEv2D<6>), GRMP_ACT_NONE, 9)                        \
  X((ReconEv2D<3>), (ReconEv2D<3>), GRMP_ACT_NONE, 3) X((ReconEv2D<6>), (ReconEv2D<6>), GRMP_ACT_NONE, 3)

#define GRMP_RECT_FORMS(X)                                                                                       \
  GRMP_RECT(X, (H1Ev<2, 2, 3, 3, GRMP_OP_DIV>), (H1Ev<2, 1, 1, 0, GRMP_OP_ID>), 1)                               \
  GRMP_RECT(X, (H1Ev<3, 3, 4, 4, GRMP_OP_DIV>), (H1Ev<3, 1, 1, 0, GRMP_OP_ID>), 4)                               \
  GRMP_RECT(X, (H1Ev<2, 2, 6, 0, GRMP_OP_DIV>), (H1Ev<2, 1, 3, 0, GRMP_OP_ID>), 3)                               \
  GRMP_RECT(X, (H1Ev<3, 3, 10, 0, GRMP_OP_DIV>), (H1Ev<3, 1, 4, 0, GRMP_OP_ID>), 4)                              \
  GRMP_RECT(X, (HdivEv<2, 3, GRMP_OP_DIV>), (H1Ev<2, 1, 1, 0, GRMP_OP_ID>), 1)                                   \
  GRMP_RECT(X, (HdivEv<2, 6, GRMP_OP_DIV>), (H1Ev<2, 1, 1, 0, GRMP_OP_ID>), 1)                                   \
  GRMP_RECT(X, (HdivEv<3, 4, GRMP_OP_DIV>), (H1Ev<3, 1, 1, 0, GRMP_OP_ID>), 1)                                   \
  GRMP_RECT(X, (HdivEv<3, 16, GRMP_OP_DIV>), (H1Ev<3, 1, 1, 0, GRMP_OP_ID>), 1)

template <class Ev> __host__ inline bool ev_matches(const ColEvalDesc& d) {
  return Ev::KIND == d.kind && Ev::ED == d.ed && Ev::NC == d.nc && Ev::NDS == d.nds && Ev::NBUB == d.nbub && Ev::OP == d.op;
}

}  // namespace grmp

// ==== closed-form ("reference tensor") evaluators ==============================================================================
// On affine cells every local entry of the forms below is
//     local[r, c] = sum_t G_t(cell) * K_t[r][c],     K_t[r][c] = sum_q w_q T_R[r][a][q] T_C[c][b][q],  t = (a, b)
// (the quadrature sum commutes with the cell's constant Jacobian; bilinearform.jl:294-317 with feevaluator_h1.jl:61-74,
// feevaluator_hdiv.jl:2-19).  K is formed ONCE from the caller's tables and weights (host, colpath_build), G once per cell
// (cell_geo_kernel), and a column thread needs NT fused multiply-adds per local row instead of a quadrature loop.
// The classes share the pair records of the quadrature evaluators above (same NROW, same row numbering).
namespace grmp {

__host__ __device__ constexpr int nsym(int ed) { return ed * (ed + 1) / 2; }
// reference tensors in shared memory: K[s_col][s][t] (t fastest), one block of cf_kblock doubles per column function, so that the
// NSF * NT numbers a column thread needs are contiguous and come in as 16-byte loads (half the shared-memory wavefronts of
// the former [t][s][s_col] layout, where every number was its own 8-byte load)
// block size = 2 (mod 4) doubles: consecutive blocks start 4 (mod 8) banks apart, so the 16-byte chunks that lanes with different s_col
// read in one instruction fall into different banks (up to 8 column functions)
__host__ __device__ constexpr int cf_kblock(int nsf, int nt) { return ((nsf * nt + 1) / 4) * 4 + 2; }
// Sc[s] = sum_t G[t] K[s_col][s][t]
template <int NSF, int NT> __device__ __forceinline__ void cf_contract(const double* __restrict__ sK, int scol, const double (&G)[NT], double (&Sc)[NSF]) {
  constexpr int B = cf_kblock(NSF, NT);
  const double2* __restrict__ kc = reinterpret_cast<const double2*>(sK + scol * B);
#pragma unroll
  for (int s = 0; s < NSF; s++) Sc[s] = 0.0;
#pragma unroll
  for (int c = 0; c < B / 2; c++) {
    const double2 kk = kc[c];
    if (2 * c < NSF * NT) Sc[(2 * c) / NT] = fma(G[(2 * c) % NT], kk.x, Sc[(2 * c) / NT]);
    if (2 * c + 1 < NSF * NT) Sc[(2 * c + 1) / NT] = fma(G[(2 * c + 1) % NT], kk.y, Sc[(2 * c + 1) / NT]);
  }
}
__host__ __device__ constexpr int sym_a(int ed, int t) { return ed == 2 ? (t == 0 ? 0 : (t == 1 ? 0 : 1)) : (t < 3 ? 0 : (t < 5 ? 1 : 2)); }
__host__ __device__ constexpr int sym_b(int ed, int t) { return ed == 2 ? (t == 0 ? 0 : (t == 1 ? 1 : 1)) : (t == 0 ? 0 : t == 1 ? 1 : t == 2 ? 2 : t == 3 ? 1 : t == 4 ? 2 : 2); }

// ---- H1 (componentwise, + Bernardi-Raugel bubbles): [Gradient, Gradient] (OPK = GRMP_OP_GRAD) or [Identity, Identity] (GRMP_OP_ID),
//      NoAction.  Scalar matrix column Sc[s] = sum_t G_t K_t[s][s_col]; entries = component weights x Sc (h1v_br.jl:150-162).
template <int ED_, int NC_, int NDS_, int NBUB_, int OPK_> struct CfH1 {
  static constexpr int ED = ED_, NC = NC_, NDS = NDS_, NBUB = NBUB_, OPK = OPK_;
  static constexpr int NSF = NDS + NBUB, NROW = NC * NDS + NBUB;
  static constexpr int NT = OPK == GRMP_OP_GRAD ? nsym(ED) : 1;      // symmetric G: K_t = K_ab + K_ba for a != b
  static constexpr bool SYMK = true;
  static constexpr int GEO_N = NT + NBUB * ED;
  static constexpr int STRIDE = (GEO_N + 1) & ~1;
  using Quad = H1Ev<ED, NC, NDS, NBUB, OPK>;                           // the quadrature evaluator with the same records / tables
  static constexpr int ACT = GRMP_ACT_NONE;
  __device__ __forceinline__ static void geo(const GridView& g, i64 cell, const double* act_p, double* out) {
    CellGeo<ED> T;
    cell_geo<ED>(g, cell, T);
    const double vol = g.vol[cell];
    if constexpr (OPK == GRMP_OP_GRAD) {
#pragma unroll
      for (int t = 0; t < NT; t++) {
        const int a = sym_a(ED, t), b = sym_b(ED, t);
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < ED; k++) s = fma(T.Ainv[k][a], T.Ainv[k][b], s);
        out[t] = vol * s;
      }
    } else out[0] = vol;
    if constexpr (NBUB > 0) {
      const i32* cf = g.cellfaces + cell * (ED + 1);
#pragma unroll
      for (int b = 0; b < NBUB; b++)
#pragma unroll
        for (int c = 0; c < ED; c++) out[NT + b * ED + c] = g.fnormals[(i64)(cf[b] - 1) * ED + c];
    }
  }
  // one (cell, column) pair: emit(row, value / factor)
  template <class F> __device__ __forceinline__ static void column(const double* __restrict__ cr, int l, const double* __restrict__ sK, F&& emit) {
    int scol;
    double beta[NC];
    if (NBUB > 0 && l >= NC * NDS) {
      const int b = l - NC * NDS;
      scol = NDS + b;
#pragma unroll
      for (int c = 0; c < NC; c++) {
        double v = 0.0;
#pragma unroll
        for (int bb = 0; bb < NBUB; bb++) v = (bb == b) ? cr[NT + bb * ED + c] : v;
        beta[c] = v;
      }
    } else {
      const int cl = l / NDS;
      scol = l - cl * NDS;
#pragma unroll
      for (int c = 0; c < NC; c++) beta[c] = (c == cl) ? 1.0 : 0.0;
    }
    double G[NT];
#pragma unroll
    for (int t = 0; t < NT; t++) G[t] = cr[t];
    double Sc[NSF];
    cf_contract<NSF, NT>(sK, scol, G, Sc);
#pragma unroll
    for (int c = 0; c < NC; c++)
#pragma unroll
      for (int s = 0; s < NDS; s++) emit(c * NDS + s, beta[c] * Sc[s]);
#pragma unroll
    for (int b = 0; b < NBUB; b++) {
      double t = 0.0;
#pragma unroll
      for (int c = 0; c < NC; c++) t = fma(cr[NT + b * ED + c], beta[c], t);
      emit(NC * NDS + b, t * Sc[NDS + b]);
    }
  }
};

// (A closed form of the Hooke form -- G[c'][c][a][b] = |T| sum_kl Ainv[k][a] E[c][k][c'][l] Ainv[l][b], 16 / 81 numbers per cell -- was
//  measured and dropped: the 128-byte geometry record per pair made it slower than the quadrature kernel on triangles (2.68 vs
//  1.52 ms at level 10), and on tetrahedra the contraction of reference tensors loses a digit on entries that are small by
//  cancellation.  Hooke forms use H1Ev<.., GRMP_OP_SYMGRAD> above.)
// ---- Hdiv [Identity, Identity] (mass, contravariant Piola, feevaluator_hdiv.jl:2-19): G_t = |T| / det^2 (A^T A)_ab, signs of the
//      reference functions from CellFaceSigns / orientations (HdivEv::negmask); rows are REFERENCE functions (records map them)
template <int ED_, int NDALL_> struct CfHdivMass {
  static constexpr int ED = ED_, NC = 1, NDS = NDALL_, NBUB = 0, OPK = GRMP_OP_ID;
  static constexpr int NSF = NDALL_, NROW = NDALL_;
  static constexpr int NT = nsym(ED);
  static constexpr bool SYMK = true;
  static constexpr int GEO_N = NT + 1;
  static constexpr int STRIDE = (GEO_N + 1) & ~1;
  using Quad = HdivEv<ED, NDALL_, GRMP_OP_ID>;
  static constexpr int ACT = GRMP_ACT_NONE;
  __device__ __forceinline__ static void geo(const GridView& g, i64 cell, const double* act_p, double* out) {
    CellGeo<ED> T;
    cell_geo<ED>(g, cell, T);
    const double f = g.vol[cell] * T.idet * T.idet;
#pragma unroll
    for (int t = 0; t < NT; t++) {
      const int a = sym_a(ED, t), b = sym_b(ED, t);
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < ED; k++) s = fma(T.A[k][a], T.A[k][b], s);
      out[t] = f * s;
    }
    out[NT] = __longlong_as_double((long long)Quad::negmask(g, cell));
  }
  template <class F> __device__ __forceinline__ static void column(const double* __restrict__ cr, int l, const double* __restrict__ sK, F&& emit) {
    const u32 neg = (u32)__double_as_longlong(cr[NT]);
    const u32 flip = ((neg >> l) & 1u) ? ~neg : neg;      // bit r: sign_r * sign_l < 0
    double G[NT];
#pragma unroll
    for (int t = 0; t < NT; t++) G[t] = cr[t];
    double Sc[NSF];
    cf_contract<NSF, NT>(sK, l, G, Sc);
#pragma unroll
    for (int r = 0; r < NDALL_; r++) emit(r, ((flip >> r) & 1u) ? -Sc[r] : Sc[r]);
  }
};

// forms with a closed-form kernel.  Measured against the quadrature kernels on one B200 (profiles/r2_all_configs.md): better for the
// P2 / Bernardi-Raugel Laplacians, all H1 mass matrices and RT0 / BDM1(2D) mass; NOT listed (quadrature kernel is faster): P1
// Laplacians (one quadrature point: nothing to save) and the BDM1 mass matrix on tetrahedra (16 x 6 table reads per pair).
#define GRMP_CF_FORMS(Y)                                                                                                   \
  Y((CfH1<2, 1, 6, 0, GRMP_OP_GRAD>)) Y((CfH1<2, 2, 6, 0, GRMP_OP_GRAD>)) Y((CfH1<2, 2, 3, 3, GRMP_OP_GRAD>))            \
  Y((CfH1<3, 1, 10, 0, GRMP_OP_GRAD>)) Y((CfH1<3, 3, 10, 0, GRMP_OP_GRAD>)) Y((CfH1<3, 3, 4, 4, GRMP_OP_GRAD>))           \
  Y((CfH1<2, 1, 3, 0, GRMP_OP_ID>)) Y((CfH1<2, 2, 3, 0, GRMP_OP_ID>)) Y((CfH1<2, 1, 6, 0, GRMP_OP_ID>))                  \
  Y((CfH1<2, 2, 6, 0, GRMP_OP_ID>)) Y((CfH1<2, 2, 3, 3, GRMP_OP_ID>))                                                     \
  Y((CfH1<3, 1, 4, 0, GRMP_OP_ID>)) Y((CfH1<3, 3, 4, 0, GRMP_OP_ID>)) Y((CfH1<3, 1, 10, 0, GRMP_OP_ID>))                 \
  Y((CfH1<3, 3, 10, 0, GRMP_OP_ID>))                                                                                      \
  Y((CfHdivMass<2, 3>)) Y((CfHdivMass<2, 6>)) Y((CfHdivMass<3, 4>))

}  // namespace grmp
