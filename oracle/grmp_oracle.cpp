// =====================================================================================
// grmp_oracle.cpp  --  TEST INFRASTRUCTURE ONLY (never linked, imported or executed by
// the product path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it).
//
// CPU restatement, operation for operation, of the assembly hot path of
// GradientRobustMultiPhysics.jl v0.12.0 (pure Julia; Julia is not available in this
// image, so the reference itself cannot be run -> **parity unpinned** at the bit level,
// see DESIGN.md).  Every function cites the reference file:line it follows (paths
// relative to /root/reference).  Compile with -ffp-contract=off: Julia does not contract
// a*b+c into FMA, and the sparsity pattern depends on exact floating-point zeros
// (src/fematrix.jl:54-58).
//
// Third-party semantics restated from their published behaviour (sources absent from
// /root/reference):  ExtendableGrids.jl >= 0.9.16 (L2GTransformer / update_trafo! /
// mapderiv!), ExtendableSparse.jl >= 1.2 (rawupdateindex! / flush!, LNK -> CSC),
// ForwardDiff.jl ^0.10.35 (dual-number jacobians of the reference bases).
// =====================================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

typedef int32_t i32;
typedef int64_t i64;

thread_local std::string g_err;

// -------------------------------------------------------------------------------------
// forward-mode dual number with 3 partials (ForwardDiff.Dual restated; product rule with
// separate multiply and add, ForwardDiff/src/dual.jl `*` -> _mul_partials(px,py,vy,vx))
// -------------------------------------------------------------------------------------
struct Dual {
  double v;
  double d[3];
  Dual() : v(0) { d[0] = d[1] = d[2] = 0; }
  Dual(double x) : v(x) { d[0] = d[1] = d[2] = 0; }
};
inline Dual operator+(const Dual& a, const Dual& b) { Dual r; r.v = a.v + b.v; for (int i = 0; i < 3; i++) r.d[i] = a.d[i] + b.d[i]; return r; }
inline Dual operator-(const Dual& a, const Dual& b) { Dual r; r.v = a.v - b.v; for (int i = 0; i < 3; i++) r.d[i] = a.d[i] - b.d[i]; return r; }
inline Dual operator-(const Dual& a) { Dual r; r.v = -a.v; for (int i = 0; i < 3; i++) r.d[i] = -a.d[i]; return r; }
inline Dual operator*(const Dual& a, const Dual& b) {
  Dual r; r.v = a.v * b.v;
  for (int i = 0; i < 3; i++) r.d[i] = (b.v * a.d[i]) + (a.v * b.d[i]);
  return r;
}
inline Dual operator*(double a, const Dual& b) { Dual r; r.v = a * b.v; for (int i = 0; i < 3; i++) r.d[i] = a * b.d[i]; return r; }
inline Dual operator*(const Dual& b, double a) { Dual r; r.v = b.v * a; for (int i = 0; i < 3; i++) r.d[i] = b.d[i] * a; return r; }
inline Dual operator-(const Dual& a, double b) { Dual r = a; r.v = a.v - b; return r; }
inline Dual operator-(double a, const Dual& b) { Dual r; r.v = a - b.v; for (int i = 0; i < 3; i++) r.d[i] = -b.d[i]; return r; }
inline Dual operator+(const Dual& a, double b) { Dual r = a; r.v = a.v + b; return r; }
inline Dual operator+(double a, const Dual& b) { Dual r = b; r.v = a + b.v; return r; }
inline Dual operator/(const Dual& a, double b) { Dual r; r.v = a.v / b; for (int i = 0; i < 3; i++) r.d[i] = a.d[i] / b; return r; }

// -------------------------------------------------------------------------------------
// enums shared with the python wrapper (oracle/oracle.py)
// -------------------------------------------------------------------------------------
enum FEType { H1P1 = 1, H1P2 = 2, H1BR = 3, HDIVRT0 = 4, HDIVBDM1 = 5, L2P0 = 6 };
enum Op { OP_ID = 1, OP_GRAD = 2, OP_SYMGRAD = 3, OP_DIV = 4, OP_RECON_ID_RT0 = 5, OP_RECON_ID_BDM1 = 6, OP_NORMALFLUX = 7 };
enum Action { ACT_NONE = 0, ACT_HOOKE2D = 1, ACT_HOOKE3D = 2, ACT_CONVECTION = 3 };
enum APT { APT_GENERAL = 0, APT_SYMMETRIC = 1, APT_LUMPED = 2 };
enum FSrc { F_NONE = 0, F_CONST = 1, F_QP_TABLE = 2 };
enum IIKind { II_NONE = 0, II_L2NORM = 1, II_L2ERROR = 2 };

inline bool is_recon(int op) { return op == OP_RECON_ID_RT0 || op == OP_RECON_ID_BDM1; }

// -------------------------------------------------------------------------------------
// reference bases (the `get_basis` closures), templated on the scalar so that the same
// literal expression gives values (double) and ForwardDiff-style jacobians (Dual).
// refbasis is (ndofs_all x ncomp), addressed rb(dof, comp) 0-based; zero-initialised by
// the caller (src/feevaluator.jl:64-68, 241-246).
// -------------------------------------------------------------------------------------
template <class T> struct RefB {
  std::vector<T> a; int nd, nc;
  RefB(int nd_, int nc_) : a((size_t)nd_ * nc_), nd(nd_), nc(nc_) {}
  T& operator()(int dof, int comp) { return a[(size_t)comp * nd + dof]; }   // column-major like Julia
  T& last() { return a.back(); }                                            // refbasis[end]
};

// src/fedefs/h1_p1.jl:64-75
template <class T> void basis_H1P1(RefB<T>& rb, const T* x, int edim, int ncomp) {
  for (int k = 1; k <= ncomp; k++) {
    int r = (edim + 1) * k - edim - 1;
    rb(r, k - 1) = T(1.0);
    for (int j = 1; j <= edim; j++) {
      rb(r, k - 1) = rb(r, k - 1) - x[j - 1];
      rb(r + j, k - 1) = x[j - 1];
    }
  }
}
// src/fedefs/h1_p2.jl:123-132 (Edge1D: the boundary faces of a 2D grid), 208-220 (Triangle2D), 223-239 (Tetrahedron3D)
template <class T> void basis_H1P2(RefB<T>& rb, const T* x, int edim, int ncomp) {
  if (edim == 1) {
    rb.last() = 1.0 - x[0];
    for (int k = 1; k <= ncomp; k++) {
      T l = rb.last();
      rb(3 * k - 3, k - 1) = 2.0 * l * (l - 0.5);
      rb(3 * k - 2, k - 1) = 2.0 * x[0] * (x[0] - 0.5);
      rb(3 * k - 1, k - 1) = 4.0 * l * x[0];
    }
  } else if (edim == 2) {
    rb.last() = 1.0 - x[0] - x[1];
    for (int k = 1; k <= ncomp; k++) {
      T l = rb.last();
      rb(6 * k - 6, k - 1) = 2.0 * l * (l - 0.5);
      rb(6 * k - 5, k - 1) = 2.0 * x[0] * (x[0] - 0.5);
      rb(6 * k - 4, k - 1) = 2.0 * x[1] * (x[1] - 0.5);
      rb(6 * k - 3, k - 1) = 4.0 * l * x[0];
      rb(6 * k - 2, k - 1) = 4.0 * x[0] * x[1];
      rb(6 * k - 1, k - 1) = 4.0 * x[1] * l;
    }
  } else {
    rb.last() = 1.0 - x[0] - x[1] - x[2];
    for (int k = 1; k <= ncomp; k++) {
      T l = rb.last();
      rb(10 * k - 10, k - 1) = 2.0 * l * (l - 0.5);
      rb(10 * k - 9, k - 1) = 2.0 * x[0] * (x[0] - 0.5);
      rb(10 * k - 8, k - 1) = 2.0 * x[1] * (x[1] - 0.5);
      rb(10 * k - 7, k - 1) = 2.0 * x[2] * (x[2] - 0.5);
      rb(10 * k - 6, k - 1) = 4.0 * l * x[0];
      rb(10 * k - 5, k - 1) = 4.0 * l * x[1];
      rb(10 * k - 4, k - 1) = 4.0 * l * x[2];
      rb(10 * k - 3, k - 1) = 4.0 * x[0] * x[1];
      rb(10 * k - 2, k - 1) = 4.0 * x[0] * x[2];
      rb(10 * k - 1, k - 1) = 4.0 * x[1] * x[2];
    }
  }
}
// src/fedefs/h1v_br.jl:117-130 (Triangle2D), 218-232 (Tetrahedron3D)
template <class T> void basis_H1BR(RefB<T>& rb, const T* x, int edim) {
  basis_H1P1(rb, x, edim, edim);
  if (edim == 2) {
    int o = 6;
    rb(o + 0, 0) = 6.0 * x[0] * rb(0, 0);
    rb(o + 1, 0) = 6.0 * x[1] * x[0];
    rb(o + 2, 0) = 6.0 * rb(0, 0) * x[1];
    for (int j = 0; j < 3; j++) rb(o + j, 1) = rb(o + j, 0);
  } else {
    int o = 12;
    rb(o + 0, 0) = 60.0 * x[0] * rb(0, 0) * x[1];
    rb(o + 1, 0) = 60.0 * rb(0, 0) * x[0] * x[2];
    rb(o + 2, 0) = 60.0 * x[0] * x[1] * x[2];
    rb(o + 3, 0) = 60.0 * rb(0, 0) * x[1] * x[2];
    for (int j = 0; j < 4; j++) for (int k = 1; k < 3; k++) rb(o + j, k) = rb(o + j, 0);
  }
}
// src/fedefs/hdiv_rt0.jl:67-73 (Triangle2D), 84-92 (Tetrahedron3D)
template <class T> void basis_RT0(RefB<T>& rb, const T* x, int edim) {
  if (edim == 2) {
    rb(0, 0) = x[0];       rb(0, 1) = x[1] - 1.0;
    rb(1, 0) = x[0];       rb(1, 1) = x[1];
    rb(2, 0) = x[0] - 1.0; rb(2, 1) = x[1];
  } else {
    rb(0, 0) = 2.0 * x[0];         rb(0, 1) = 2.0 * x[1];         rb(0, 2) = 2.0 * (x[2] - 1.0);
    rb(1, 0) = 2.0 * x[0];         rb(1, 1) = 2.0 * (x[1] - 1.0); rb(1, 2) = 2.0 * x[2];
    rb(2, 0) = 2.0 * x[0];         rb(2, 1) = 2.0 * x[1];         rb(2, 2) = 2.0 * x[2];
    rb(3, 0) = 2.0 * (x[0] - 1.0); rb(3, 1) = 2.0 * x[1];         rb(3, 2) = 2.0 * x[2];
  }
}
// src/fedefs/hdiv_bdm1.jl:82-93 (Triangle2D), 119-165 (Tetrahedron3D, 16 functions)
template <class T> void basis_BDM1(RefB<T>& rb, const T* x, int edim) {
  if (edim == 2) {
    rb(0, 0) = x[0];       rb(0, 1) = x[1] - 1.0;
    rb(2, 0) = x[0];       rb(2, 1) = x[1];
    rb(4, 0) = x[0] - 1.0; rb(4, 1) = x[1];
    rb(1, 0) = 6.0 * x[0];                           rb(1, 1) = 6.0 - 12.0 * x[0] - 6.0 * x[1];
    rb(3, 0) = -6.0 * x[0];                          rb(3, 1) = 6.0 * x[1];
    rb(5, 0) = 6.0 * (x[0] - 1.0) + 12.0 * x[1];     rb(5, 1) = -6.0 * x[1];
  } else {
    rb(0, 0) = 2.0 * x[0];          rb(0, 1) = 2.0 * x[1];          rb(0, 2) = 2.0 * (x[2] - 1.0);
    rb(4, 0) = 2.0 * x[0];          rb(4, 1) = 2.0 * (x[1] - 1.0);  rb(4, 2) = 2.0 * x[2];
    rb(8, 0) = 2.0 * x[0];          rb(8, 1) = 2.0 * x[1];          rb(8, 2) = 2.0 * x[2];
    rb(12, 0) = 2.0 * (x[0] - 1.0); rb(12, 1) = 2.0 * x[1];         rb(12, 2) = 2.0 * x[2];
    rb.last() = 1.0 - x[0] - x[1] - x[2];
    T l = rb.last();
    T zero(0.0);
    // face 1
    rb(1, 0) = 24.0 * x[0];   rb(1, 1) = zero;            rb(1, 2) = 24.0 * (l - x[0]);
    rb(2, 0) = zero;          rb(2, 1) = -24.0 * x[1];    rb(2, 2) = -24.0 * (l - x[1]);
    rb(3, 0) = -24.0 * x[0];  rb(3, 1) = 24.0 * x[1];     rb(3, 2) = -24.0 * (x[1] - x[0]);
    // face 2
    rb(5, 0) = zero;          rb(5, 1) = 24.0 * (l - x[2]);    rb(5, 2) = 24.0 * x[2];
    rb(6, 0) = -24.0 * x[0];  rb(6, 1) = -24.0 * (l - x[0]);   rb(6, 2) = zero;
    rb(7, 0) = 24.0 * x[0];   rb(7, 1) = -24.0 * (x[0] - x[2]); rb(7, 2) = -24.0 * x[2];
    // face 3
    rb(9, 0) = -24.0 * x[0];  rb(9, 1) = zero;            rb(9, 2) = 24.0 * x[2];
    rb(10, 0) = 24.0 * x[0];  rb(10, 1) = -24.0 * x[1];   rb(10, 2) = zero;
    rb(11, 0) = zero;         rb(11, 1) = 24.0 * x[1];    rb(11, 2) = -24.0 * x[2];
    // face 4 (the last assignment overwrites refbasis[end] == rb(15,2), as in the reference)
    rb(13, 0) = 24.0 * (l - x[1]);   rb(13, 1) = 24.0 * x[1];   rb(13, 2) = zero;
    rb(14, 0) = -24.0 * (l - x[2]);  rb(14, 1) = zero;          rb(14, 2) = -24.0 * x[2];
    rb(15, 0) = -24.0 * (x[2] - x[1]); rb(15, 1) = -24.0 * x[1]; rb(15, 2) = 24.0 * x[2];
  }
}
// src/fedefs/l2_p0.jl (constant 1 per component)
template <class T> void basis_L2P0(RefB<T>& rb, const T*, int ncomp) {
  for (int k = 0; k < ncomp; k++) rb(k, k) = T(1.0);
}

struct FEInfo { int ncomp, nd, nd_all, polyorder; bool coeffs, hdiv; };

// get_ndofs / get_ndofs_all / get_polynomialorder (src/fedefs/*.jl headers)
FEInfo fe_info(int fe, int ncomp, int edim, bool on_faces = false) {
  FEInfo r{};
  r.ncomp = ncomp; r.coeffs = false; r.hdiv = false;
  int nn = edim + 1, nf = edim + 1, ne = (edim == 1) ? 1 : (edim == 2) ? 3 : 6;   // Edge1D: "N1I1" (h1_p2.jl:37)
  switch (fe) {
    case H1P1: r.nd = r.nd_all = nn * ncomp; r.polyorder = 1; break;
    case H1P2: r.nd = r.nd_all = (nn + ne) * ncomp; r.polyorder = 2; break;
    case H1BR: r.ncomp = edim; r.nd = r.nd_all = nf + nn * edim; r.polyorder = (edim == 2) ? 2 : 3; r.coeffs = true; break;
    case HDIVRT0: r.ncomp = edim; r.nd = r.nd_all = nf; r.polyorder = 1; r.hdiv = true; break;
    case HDIVBDM1: r.ncomp = edim; r.nd = edim * nf; r.nd_all = (edim == 2) ? 2 * nf : 4 * nf; r.polyorder = 1; r.hdiv = true; break;
    case L2P0: r.nd = r.nd_all = ncomp; r.polyorder = 0; break;
    default: r.nd = -1;
  }
  if (on_faces && (fe == HDIVRT0 || fe == HDIVBDM1)) {       // normal-flux face bases: get_ndofs / get_polynomialorder on the face geometry
    r.ncomp = 1; r.nd = r.nd_all = (fe == HDIVRT0) ? 1 : edim + 1;      // hdiv_rt0.jl:21, 24, 26; hdiv_bdm1.jl:20-21, 26, 29
    r.polyorder = (fe == HDIVRT0) ? 0 : 1;
  }
  return r;
}

// normal-flux bases on faces: hdiv_rt0.jl:61-65, hdiv_bdm1.jl:74-79 (Edge1D), 109-115 (Triangle2D); one component
template <class T> void basis_hdiv_face(int fe, RefB<T>& rb, const T* x, int fdim) {
  rb(0, 0) = T(1.0);
  if (fe == HDIVBDM1 && fdim == 1) rb(1, 0) = 12.0 * (x[0] - 0.5);
  if (fe == HDIVBDM1 && fdim == 2) {
    rb(1, 0) = 12.0 * (2.0 * x[0] + x[1] - 1.0);
    rb(2, 0) = 12.0 * (2.0 * x[1] + x[0] - 1.0);
  }
}
template <class T> void eval_basis(int fe, RefB<T>& rb, const T* x, int edim, int ncomp) {
  switch (fe) {
    case H1P1: basis_H1P1(rb, x, edim, ncomp); break;
    case H1P2: basis_H1P2(rb, x, edim, ncomp); break;
    case H1BR: basis_H1BR(rb, x, edim); break;
    case HDIVRT0: basis_RT0(rb, x, edim); break;
    case HDIVBDM1: basis_BDM1(rb, x, edim); break;
    case L2P0: basis_L2P0(rb, x, ncomp); break;
  }
}

// -------------------------------------------------------------------------------------
// quadrature rules (src/quadrature.jl:173-195 triangle, 268-325 tetrahedron,
// 332-502 symmetric rules, 528-562 Stroud conical product)
// -------------------------------------------------------------------------------------
struct QRule { int dim; std::vector<double> xref; std::vector<double> w; int n() const { return (int)w.size(); } };

// cyclic Jacobi eigen-decomposition of a small symmetric matrix; eigenvalues ascending,
// eigenvectors normalised (stand-in for LinearAlgebra.eigen, quadrature.jl:533,540)
void sym_eigen(int n, std::vector<double> a, std::vector<double>& vals, std::vector<double>& vecs) {
  std::vector<double> v((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) v[i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += a[p * n + q] * a[p * n + q];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
      if (a[p * n + q] == 0.0) continue;
      double theta = (a[q * n + q] - a[p * n + p]) / (2 * a[p * n + q]);
      double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
      double c = 1 / std::sqrt(t * t + 1), s = t * c;
      for (int k = 0; k < n; k++) {
        double akp = a[k * n + p], akq = a[k * n + q];
        a[k * n + p] = c * akp - s * akq; a[k * n + q] = s * akp + c * akq;
      }
      for (int k = 0; k < n; k++) {
        double apk = a[p * n + k], aqk = a[q * n + k];
        a[p * n + k] = c * apk - s * aqk; a[q * n + k] = s * apk + c * aqk;
      }
      for (int k = 0; k < n; k++) {
        double vkp = v[k * n + p], vkq = v[k * n + q];
        v[k * n + p] = c * vkp - s * vkq; v[k * n + q] = s * vkp + c * vkq;
      }
    }
  }
  std::vector<int> idx(n);
  for (int i = 0; i < n; i++) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int x, int y) { return a[x * n + x] < a[y * n + y]; });
  vals.resize(n); vecs.assign((size_t)n * n, 0.0);
  for (int j = 0; j < n; j++) {
    vals[j] = a[idx[j] * n + idx[j]];
    for (int k = 0; k < n; k++) vecs[k * n + j] = v[k * n + idx[j]];   // vecs(k,j): component k of eigenvector j
  }
}

// quadrature.jl:528-562
QRule stroud_rule(int order) {
  int n = order / 2 + 1;
  std::vector<double> A((size_t)n * n, 0.0), r, a, s, b, vec;
  for (int k = 1; k <= n - 1; k++) {
    double g = k / std::sqrt(4.0 * k * k - 1.0);
    A[(k - 1) * n + k] = g; A[k * n + (k - 1)] = g;
  }
  sym_eigen(n, A, r, vec);
  a.resize(n);
  for (int j = 0; j < n; j++) a[j] = 2 * vec[0 * n + j] * vec[0 * n + j];
  std::fill(A.begin(), A.end(), 0.0);
  for (int k = 1; k <= n; k++) A[(k - 1) * n + (k - 1)] = -1.0 / (4.0 * k * k - 1.0);
  for (int k = 1; k <= n - 1; k++) {
    double g = std::sqrt((double)(k + 1) * k) / (2.0 * (k + 1) - 1.0);
    A[(k - 1) * n + k] = g; A[k * n + (k - 1)] = g;
  }
  sym_eigen(n, A, s, vec);
  b.resize(n);
  for (int j = 0; j < n; j++) b[j] = 2 * vec[0 * n + j] * vec[0 * n + j];
  for (int j = 0; j < n; j++) { r[j] = .5 * r[j] + .5; s[j] = .5 * s[j] + .5; a[j] = .5 * a[j]; b[j] = .5 * b[j]; }
  QRule q; q.dim = 2;
  for (int js = 0; js < n; js++) for (int ir = 0; ir < n; ir++) {
    q.xref.push_back(s[js] * 1.0 - (r[ir] * (s[js] - 1)) * 0.0);
    q.xref.push_back(s[js] * 0.0 - (r[ir] * (s[js] - 1)) * 1.0);
    q.w.push_back(a[ir] * b[js]);
  }
  return q;
}

// quadrature.jl:334-410 (order-8 branch; Zhang/Cui/Liu 2009)
QRule symmetric_rule_tri8() {
  const double wS3 = .1443156076777871682510911104890646;
  const double aS21[3] = {.1705693077517602066222935014914645, .0505472283170309754584235505965989, .4592925882927231560288155144941693};
  const double wS21[3] = {.1032173705347182502817915502921290, .0324584976231980803109259283417806, .0950916342672846247938961043885843};
  const double aS111[2] = {.2631128296346381134217857862846436, .0083947774099576053372138345392944};
  const double wS111 = .0272303141744349942648446900739089;
  QRule q; q.dim = 2;
  auto add = [&](double x, double y, double w) { q.xref.push_back(x); q.xref.push_back(y); q.w.push_back(w); };
  add(1.0 / 3, 1.0 / 3, wS3);
  for (int j = 0; j < 3; j++) {
    add(aS21[j], aS21[j], wS21[j]); add(aS21[j], 1 - 2 * aS21[j], wS21[j]); add(1 - 2 * aS21[j], aS21[j], wS21[j]);
  }
  double a = aS111[0], b = aS111[1];
  add(a, b, wS111); add(b, a, wS111); add(a, 1 - a - b, wS111); add(b, 1 - a - b, wS111); add(1 - a - b, a, wS111); add(1 - a - b, b, wS111);
  return q;
}
// quadrature.jl:417-502 (order <= 8, 46 points)
QRule symmetric_rule_tet8() {
  const double aS31[4] = {.0396754230703899012650713295393895, .3144878006980963137841605626971483, .1019866930627033000000000000000000, .1842036969491915122759464173489092};
  const double wS31[4] = {.0063971477799023213214514203351730, .0401904480209661724881611584798178, .0243079755047703211748691087719226, .0548588924136974404669241239903914};
  const double aS22 = .0634362877545398924051412387018983, wS22 = .0357196122340991824649509689966176;
  const double aS211[2][2] = {{.0216901620677280048026624826249302, .7199319220394659358894349533527348}, {.2044800806367957142413355748727453, .5805771901288092241753981713906204}};
  const double wS211[2] = {.0071831906978525394094511052198038, .0163721819453191175409381397561191};
  QRule q; q.dim = 3;
  auto add = [&](double x, double y, double z, double w) { q.xref.push_back(x); q.xref.push_back(y); q.xref.push_back(z); q.w.push_back(w); };
  for (int j = 0; j < 4; j++) {
    double a = aS31[j], c = 1 - 3 * a;
    add(a, a, a, wS31[j]); add(a, a, c, wS31[j]); add(a, c, a, wS31[j]); add(c, a, a, wS31[j]);
  }
  {
    double a = aS22, h = 0.5 - aS22;
    add(a, a, h, wS22); add(a, h, a, wS22); add(h, a, a, wS22); add(h, a, h, wS22); add(h, h, a, wS22); add(a, h, h, wS22);
  }
  for (int j = 0; j < 2; j++) {
    double a = aS211[j][0], b = aS211[j][1], c = 1 - 2 * a - b, w = wS211[j];
    add(a, a, b, w); add(a, b, a, w); add(b, a, a, w); add(a, a, c, w); add(a, c, a, w); add(c, a, a, w);
    add(a, b, c, w); add(a, c, b, w); add(c, a, b, w); add(b, a, c, w); add(b, c, a, w); add(c, b, a, w);
  }
  return q;
}

// optional caller-supplied rule (tests hand the host mirror's LAPACK-generated Stroud points to the
// oracle so that GPU-vs-oracle comparisons on eigen-generated rules can be bitwise)
QRule g_override; int g_override_edim = -1, g_override_order = -1;

bool make_qrule(int edim, int order, QRule& q) {
  if (edim == g_override_edim && order == g_override_order) { q = g_override; return true; }
  q = QRule(); q.dim = edim;
  if (edim == 1) {                                          // quadrature.jl:130-148, Gauss rule 506-525
    if (order <= 1) { q.xref = {0.5}; q.w = {1.0}; }
    else if (order == 2) { q.xref = {0.0, 0.5, 1.0}; q.w = {1.0 / 6, 2.0 / 3, 1.0 / 6}; }
    else {
      int n = order / 2 + 1;
      std::vector<double> A((size_t)n * n, 0.0), r, vec;
      for (int k = 1; k <= n - 1; k++) {
        double g = k / std::sqrt(4.0 * k * k - 1.0);
        A[(k - 1) * n + k] = g; A[k * n + (k - 1)] = g;
      }
      sym_eigen(n, A, r, vec);
      for (int j = 0; j < n; j++) { q.xref.push_back(.5 * r[j] + .5); q.w.push_back(.5 * (2 * vec[0 * n + j] * vec[0 * n + j])); }
    }
  } else if (edim == 2) {                                          // quadrature.jl:173-195
    if (order <= 1) { q.xref = {1.0 / 3, 1.0 / 3}; q.w = {1.0}; }
    else if (order == 2) { q.xref = {0.5, 0.5, 0.0, 0.5, 0.5, 0.0}; q.w = {1.0 / 3, 1.0 / 3, 1.0 / 3}; }
    else if (order == 8) q = symmetric_rule_tri8();
    else if (order <= 11) q = stroud_rule(order);
    else { g_err = "triangle quadrature order > 11 not restated"; return false; }
  } else if (edim == 3) {                                   // quadrature.jl:268-325
    if (order <= 1) { q.xref = {0.25, 0.25, 0.25}; q.w = {1.0}; }
    else if (order == 2) {
      const double a = 0.1381966011250105, b = 0.5854101966249685;
      q.xref = {a, a, a, b, a, a, a, b, a, a, a, b}; q.w = {0.25, 0.25, 0.25, 0.25};
    } else if (order <= 3) {
      q.xref = {1.0 / 4, 1.0 / 4, 1.0 / 4, 1.0 / 2, 1.0 / 6, 1.0 / 6, 1.0 / 6, 1.0 / 6, 1.0 / 6, 1.0 / 6, 1.0 / 6, 1.0 / 2, 1.0 / 6, 1.0 / 2, 1.0 / 6};
      q.w = {-4.0 / 5, 9.0 / 20, 9.0 / 20, 9.0 / 20, 9.0 / 20};
    } else if (order <= 4) {
      const double c = 0.2500000000000000, d = 0.7857142857142857, e = 0.0714285714285714, f = 0.1005964238332008, g = 0.3994035761667992;
      q.xref = {c, c, c, d, e, e, e, e, e, e, e, d, e, d, e, f, g, g, g, f, g, g, g, f, g, f, f, f, g, f, f, f, g};
      const double w0 = -0.0789333333333333, w1 = 0.0457333333333333, w2 = 0.1493333333333333;
      q.w = {w0, w1, w1, w1, w1, w2, w2, w2, w2, w2, w2};
    } else q = symmetric_rule_tet8();
  } else { g_err = "unsupported dimension"; return false; }
  return true;
}

// -------------------------------------------------------------------------------------
// grid / space views (1-based Int32 arrays laid out like the Julia column-major arrays)
// -------------------------------------------------------------------------------------
struct Grid {
  int dim, xdim; i64 nnodes, ncells, nfaces;      // xdim > dim: the items are boundary faces (AT = ON_BFACES), no affine inverse
  const double* coords; const i32* cellnodes; const double* vol; const i32* regions;
  const i32* cellfaces; const i32* signs; const i32* orient; const double* fnormals; const double* fvol;
};
struct Space { int fe, ncomp; i64 ndofs; int nd; const i32* celldofs; };

// L2GTransformer for simplices (ExtendableGrids semantics; call sites
// src/feevaluator.jl:371-390): A[:,j] = x_{j+1}-x_1, b = x_1; Ainv = A^{-T}.
struct Trafo {
  int d; double A[3][3], Ainv[3][3], b[3], det;
  void update(const Grid& g, i64 cell) {                    // update_trafo!
    d = g.dim;
    const i32* cn = g.cellnodes + cell * (d + 1);
    const double* x0 = g.coords + (i64)(cn[0] - 1) * d;
    for (int k = 0; k < d; k++) b[k] = x0[k];
    for (int j = 0; j < d; j++) {
      const double* xj = g.coords + (i64)(cn[j + 1] - 1) * d;
      for (int k = 0; k < d; k++) A[k][j] = xj[k] - b[k];
    }
    if (d == 2) det = A[0][0] * A[1][1] - A[0][1] * A[1][0];
    else det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) + A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
  }
  void mapderiv(const Grid& g, i64 cell) {                   // mapderiv! : det = d! * |T|
    if (d == 2) {
      double dt = 2 * g.vol[cell];
      Ainv[1][1] = A[0][0] / dt; Ainv[1][0] = -A[0][1] / dt; Ainv[0][1] = -A[1][0] / dt; Ainv[0][0] = A[1][1] / dt;
    } else {
      double dt = 6 * g.vol[cell];
      Ainv[0][0] = (A[1][1] * A[2][2] - A[1][2] * A[2][1]) / dt;
      Ainv[0][1] = -(A[1][0] * A[2][2] - A[1][2] * A[2][0]) / dt;
      Ainv[0][2] = (A[1][0] * A[2][1] - A[1][1] * A[2][0]) / dt;
      Ainv[1][0] = -(A[0][1] * A[2][2] - A[0][2] * A[2][1]) / dt;
      Ainv[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) / dt;
      Ainv[1][2] = -(A[0][0] * A[2][1] - A[0][1] * A[2][0]) / dt;
      Ainv[2][0] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) / dt;
      Ainv[2][1] = -(A[0][0] * A[1][2] - A[0][2] * A[1][0]) / dt;
      Ainv[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) / dt;
    }
  }
};

const int TRI_FACE[3][2] = {{0, 1}, {1, 2}, {2, 0}};
const int TET_FACE[4][3] = {{0, 2, 1}, {0, 1, 3}, {1, 2, 3}, {0, 3, 2}};

// -------------------------------------------------------------------------------------
// FEEvaluator (src/feevaluator.jl:34-138, reconstruction constructor 142-217)
// -------------------------------------------------------------------------------------
struct Evaluator {
  const Grid* g; Space sp; int op; int edim, ncomp, nd, nd_all, resultdim, nq;
  bool coeffs_flag, hdiv;
  FEInfo fi;
  // reconstruction
  int rfe = 0; FEInfo ri{}; int nd2 = 0, nd2_all = 0;
  std::vector<double> refvals;    // [i][dof_all][comp]   (refbasisvals[i][dof,comp]); for recon: of the Hdiv space
  std::vector<double> refderiv;   // [i][j][row]  row = dof + comp*nd_all (refbasisderivvals[row,j,i])
  std::vector<double> cvals;      // [i][dof][k]  (cvals[k,dof,i])
  std::vector<double> coeff;      // [dof][k]     (coefficients[k,dof])
  std::vector<double> rcoeff;     // [dofR][dofBR] (coefficients2[dofBR,dofR]); NOT zeroed between cells
  std::vector<double> tempeval;   // [i][dofR_all][k]
  std::vector<int> subset;
  std::vector<int> compress;
  Trafo T;
  i64 citem = -1;

  double& cv(int k, int dof, int i) { return cvals[((size_t)i * nd + dof) * resultdim + k]; }
  double rv(int i, int dof, int c) const { int nda = rfe ? nd2_all : nd_all; int nc = rfe ? ri.ncomp : ncomp; return refvals[((size_t)i * nda + dof) * nc + c]; }
  double rd(int row, int j, int i) const { return refderiv[((size_t)i * edim + j) * (nd_all * ncomp) + row]; }
  double& co(int k, int dof) { return coeff[(size_t)dof * ncomp + k]; }
  double& rc(int dofBR, int dofR) { return rcoeff[(size_t)dofR * nd + dofBR]; }
  double& te(int k, int dof, int i) { return tempeval[((size_t)i * nd2_all + dof) * ncomp + k]; }

  bool init(const Grid* g_, const Space& sp_, int op_, const QRule& q) {
    g = g_; sp = sp_; op = op_; edim = g->dim; nq = q.n();
    const bool on_faces = g->xdim > g->dim;
    fi = fe_info(sp.fe, sp.ncomp, edim, on_faces);
    if (fi.nd < 0) { g_err = "unknown FEType"; return false; }
    if (fi.nd != sp.nd) { g_err = "celldofs width does not match FEType"; return false; }
    ncomp = fi.ncomp; nd = fi.nd; nd_all = fi.nd_all; coeffs_flag = fi.coeffs; hdiv = fi.hdiv;
    // Length4Operator (src/functionoperators.jl:260-278)
    switch (op) {
      case OP_ID: case OP_RECON_ID_RT0: case OP_RECON_ID_BDM1: resultdim = ncomp; break;
      case OP_GRAD: resultdim = edim * ncomp; break;
      case OP_SYMGRAD: resultdim = ((edim == 2) ? 3 : 6) * ((ncomp + edim - 1) / edim); break;
      case OP_DIV: resultdim = (ncomp + edim - 1) / edim; break;
      case OP_NORMALFLUX: resultdim = 1; break;
      default: g_err = "unknown operator"; return false;
    }
    if ((op == OP_NORMALFLUX) != (on_faces && hdiv)) { g_err = "NormalFlux: Hdiv elements on boundary faces, and nothing else of them there"; return false; }
    if (op == OP_SYMGRAD && ncomp != edim) { g_err = "SymmetricGradient needs ncomponents == dim"; return false; }
    if (hdiv && !on_faces && !(op == OP_ID || op == OP_DIV)) { g_err = "Hdiv elements: Identity/Divergence only"; return false; }
    if (on_faces && !hdiv && !(op == OP_ID && (sp.fe == H1P1 || sp.fe == H1P2))) { g_err = "boundary faces: Identity of H1P1 / H1P2 only"; return false; }
    if (sp.fe == L2P0 && op != OP_ID) { g_err = "L2P0: Identity only"; return false; }
    if (is_recon(op) && sp.fe != H1BR) { g_err = "ReconstructionIdentity restated for H1BR only"; return false; }
    if ((hdiv || sp.fe == H1BR) && !on_faces && !(g->cellfaces && g->fnormals)) { g_err = "face data missing on grid"; return false; }
    cvals.assign((size_t)nq * nd * resultdim, 0.0);
    subset.resize(std::max(nd, 16)); for (size_t k = 0; k < subset.size(); k++) subset[k] = (int)k;
    if (coeffs_flag || hdiv) coeff.assign((size_t)nd * ncomp, 1.0);
    if (op == OP_SYMGRAD) compress = (edim == 2) ? std::vector<int>{1, 3, 3, 2} : std::vector<int>{1, 6, 5, 6, 2, 4, 5, 4, 3};

    int eval_fe = sp.fe, eval_nc = sp.ncomp;
    if (is_recon(op)) {
      rfe = (op == OP_RECON_ID_RT0) ? HDIVRT0 : HDIVBDM1;
      ri = fe_info(rfe, edim, edim); nd2 = ri.nd; nd2_all = ri.nd_all;
      eval_fe = rfe; eval_nc = edim;
      coeff.assign((size_t)nd2 * ncomp, 1.0);
      rcoeff.assign((size_t)nd * nd2, 0.0);
      tempeval.assign((size_t)nq * nd2_all * ncomp, 0.0);
    }
    int nda = rfe ? nd2_all : nd_all, nc = rfe ? ri.ncomp : ncomp;
    // reference values (feevaluator.jl:64-68, 95-98; recon 171-176)
    refvals.assign((size_t)nq * nda * nc, 0.0);
    for (int i = 0; i < nq; i++) {
      RefB<double> rb(nda, nc);
      if (op == OP_NORMALFLUX) basis_hdiv_face<double>(sp.fe, rb, &q.xref[(size_t)i * edim], edim);
      else eval_basis<double>(eval_fe, rb, &q.xref[(size_t)i * edim], edim, eval_nc);
      for (int dof = 0; dof < nda; dof++) for (int c = 0; c < nc; c++) refvals[((size_t)i * nda + dof) * nc + c] = rb(dof, c);
    }
    // reference derivatives via dual numbers (feevaluator.jl:112-119, 235-293)
    if (op == OP_GRAD || op == OP_SYMGRAD || op == OP_DIV) {
      refderiv.assign((size_t)nq * edim * nd_all * ncomp, 0.0);
      for (int i = 0; i < nq; i++) {
        Dual x[3];
        for (int j = 0; j < edim; j++) { x[j] = Dual(q.xref[(size_t)i * edim + j]); x[j].d[j] = 1.0; }
        RefB<Dual> rb(nd_all, ncomp);
        eval_basis<Dual>(sp.fe, rb, x, edim, sp.ncomp);
        for (int c = 0; c < ncomp; c++) for (int dof = 0; dof < nd_all; dof++) for (int j = 0; j < edim; j++)
          refderiv[((size_t)i * edim + j) * (nd_all * ncomp) + dof + c * nd_all] = rb(dof, c).d[j];
      }
    }
    // Identity of plain H1 / L2 elements is cell-independent (feevaluator.jl:100-103)
    if (op == OP_ID && !coeffs_flag && !hdiv)
      for (int i = 0; i < nq; i++) for (int j = 0; j < nd; j++) for (int k = 0; k < ncomp; k++) cv(k, j, i) = rv(i, j, k);
    return true;
  }

  // get_coefficients closures
  void update_coefficients(i64 cell, int fe, int ndc) {
    int nf = edim + 1;
    const i32* sg = g->signs ? g->signs + cell * nf : nullptr;
    if (fe == H1BR) {                                       // h1v_br.jl:150-162, 253-273
      std::fill(coeff.begin(), coeff.end(), 1.0);
      const i32* cf = g->cellfaces + cell * nf;
      for (int f = 0; f < nf; f++) for (int k = 0; k < edim; k++) co(k, edim * nf + f) = g->fnormals[(i64)(cf[f] - 1) * edim + k];
    } else if (fe == HDIVRT0) {                             // hdiv_rt0.jl:106-116
      for (int j = 0; j < nf; j++) for (int k = 0; k < ncomp; k++) co(k, j) = (double)sg[j];
    } else if (fe == HDIVBDM1 && edim == 2) {               // hdiv_bdm1.jl (2D coefficients)
      std::fill(coeff.begin(), coeff.begin() + (size_t)ndc * ncomp, 1.0);
      for (int j = 0; j < nf; j++) for (int k = 0; k < edim; k++) co(k, 2 * j) = (double)sg[j];
    } else if (fe == HDIVBDM1) {                            // hdiv_bdm1.jl (3D coefficients)
      std::fill(coeff.begin(), coeff.begin() + (size_t)ndc * ncomp, 1.0);
      for (int j = 0; j < nf; j++) for (int k = 0; k < edim; k++) { co(k, 3 * j) = (double)sg[j]; co(k, 3 * j + 1) = -1.0; co(k, 3 * j + 2) = 1.0; }
    }
  }
  // get_basissubset (hdiv_bdm1.jl, 3D): shift4orientation1 = [1,0,1,2], shift4orientation2 = [2,2,0,1]
  void update_subset(i64 cell, int fe) {
    if (fe == HDIVBDM1 && edim == 3) {
      static const int s1[4] = {1, 0, 1, 2}, s2[4] = {2, 2, 0, 1};
      const i32* o = g->orient + cell * 4;
      for (int j = 1; j <= 4; j++) {
        subset[3 * j - 3] = 4 * j - 3 - 1;
        subset[3 * j - 2] = 4 * j - s1[o[j - 1] - 1] - 1;
        subset[3 * j - 1] = 4 * j - s2[o[j - 1] - 1] - 1;
      }
    }
  }
  // boundary_coefficients! (src/reconstructions.jl:353-403 2D, 474-535 3D)
  void update_rcoeffs(i64 cell) {
    int nf = edim + 1;
    const i32* cf = g->cellfaces + cell * nf;
    if (edim == 2) {
      for (int f = 0; f < 3; f++) {
        i64 face = cf[f] - 1;
        double fv = g->fvol[face];
        for (int n = 0; n < 2; n++) {
          int node = TRI_FACE[f][n];
          for (int k = 0; k < 2; k++) {
            double nk = g->fnormals[face * 2 + k];
            if (op == OP_RECON_ID_RT0) rc(3 * k + node, f) = 0.5 * fv * nk;
            else {
              rc(3 * k + node, 2 * f) = 0.5 * fv * nk;
              double c12 = (n == 0) ? (-1.0 / 12) : (1.0 / 12);
              rc(3 * k + node, 2 * f + 1) = c12 * fv * nk * (double)g->signs[cell * 3 + f];
            }
          }
        }
        if (op == OP_RECON_ID_RT0) rc(6 + f, f) = fv; else rc(6 + f, 2 * f) = fv;
      }
    } else {
      static const double B[3][3] = {{-1.0 / 36, -1.0 / 36, 1.0 / 18}, {-1.0 / 36, 1.0 / 18, -1.0 / 36}, {1.0 / 18, -1.0 / 36, -1.0 / 36}};
      static const int r1[4] = {2, 2, 3, 1}, r2[4] = {1, 3, 1, 2};
      for (int f = 0; f < 4; f++) {
        i64 face = cf[f] - 1;
        double fv = g->fvol[face];
        for (int k = 0; k < 3; k++) {
          double nk = g->fnormals[face * 3 + k];
          for (int n = 0; n < 3; n++) {
            int node = TET_FACE[f][n];
            if (op == OP_RECON_ID_RT0) rc(4 * k + node, f) = (1.0 / 3) * fv * nk;
            else {
              int o = g->orient[cell * 4 + f] - 1;
              rc(4 * k + node, 3 * f) = (1.0 / 3) * nk * fv;
              rc(4 * k + node, 3 * f + 1) = B[n][r1[o] - 1] * nk * fv;
              rc(4 * k + node, 3 * f + 2) = B[n][r2[o] - 1] * nk * fv;
            }
          }
        }
        if (op == OP_RECON_ID_RT0) rc(12 + f, f) = fv; else rc(12 + f, 3 * f) = fv;
      }
    }
  }

  // update_basis! dispatch (src/feevaluator_h1.jl, src/feevaluator_hdiv.jl)
  void update(i64 cell) {
    if (citem == cell) return;
    citem = cell;
    if (is_recon(op)) {                                     // feevaluator_h1.jl:342-381
      T.update(*g, cell);
      update_coefficients(cell, rfe, nd2);
      update_subset(cell, rfe);
      double det = T.det;
      std::fill(tempeval.begin(), tempeval.end(), 0.0);
      for (int i = 0; i < nq; i++) for (int dof = 0; dof < nd2; dof++) for (int k = 0; k < ncomp; k++) {
        for (int l = 0; l < ncomp; l++) te(k, dof, i) += T.A[k][l] * rv(i, subset[dof], l);
        te(k, dof, i) *= co(k, dof) / det;
      }
      update_rcoeffs(cell);
      std::fill(cvals.begin(), cvals.end(), 0.0);
      for (int di = 0; di < nd; di++) for (int dj = 0; dj < nd2; dj++) {
        if (rc(di, dj) != 0)
          for (int i = 0; i < nq; i++) for (int k = 0; k < ncomp; k++) cv(k, di, i) += rc(di, dj) * te(k, dj, i);
      }
      return;
    }
    if (op == OP_NORMALFLUX) {                              // feevaluator_hdiv.jl:42-50
      for (int i = 0; i < nq; i++) for (int dof = 0; dof < nd; dof++) for (int k = 0; k < resultdim; k++) cv(k, dof, i) = rv(i, dof, k) / g->vol[cell];
      return;
    }
    if (hdiv) {
      T.update(*g, cell);
      update_subset(cell, sp.fe);
      update_coefficients(cell, sp.fe, nd);
      double det = T.det;
      std::fill(cvals.begin(), cvals.end(), 0.0);
      if (op == OP_ID) {                                    // feevaluator_hdiv.jl:2-19
        for (int i = 0; i < nq; i++) for (int dof = 0; dof < nd; dof++) for (int k = 0; k < edim; k++) {
          for (int l = 0; l < edim; l++) cv(k, dof, i) += T.A[k][l] * rv(i, subset[dof], l);
          cv(k, dof, i) *= co(k, dof) / det;
        }
      } else {                                              // OP_DIV, feevaluator_hdiv.jl:54-71
        for (int i = 0; i < nq; i++) for (int dof = 0; dof < nd; dof++) {
          for (int j = 0; j < edim; j++) cv(0, dof, i) += rd(subset[dof] + j * nd_all, j, i);
          cv(0, dof, i) *= co(0, dof) / det;
        }
      }
      return;
    }
    // H1 / L2 elements
    if (op == OP_ID) {
      if (coeffs_flag) {                                    // feevaluator_h1.jl:33-43
        update_coefficients(cell, sp.fe, nd);
        std::fill(cvals.begin(), cvals.end(), 0.0);
        for (int i = 0; i < nq; i++) for (int dof = 0; dof < nd; dof++) for (int k = 0; k < ncomp; k++) cv(k, dof, i) = rv(i, dof, k) * co(k, dof);
      }
      return;                                               // plain H1: nothing to do (feevaluator_h1.jl:2-14)
    }
    T.update(*g, cell); T.mapderiv(*g, cell);               // _update_trafo! (feevaluator.jl:371-380)
    if (coeffs_flag) update_coefficients(cell, sp.fe, nd);
    std::fill(cvals.begin(), cvals.end(), 0.0);
    if (op == OP_GRAD && !coeffs_flag) {                    // feevaluator_h1.jl:61-74, loop nest i,c,j,k,dof
      for (int i = 0; i < nq; i++) for (int c = 0; c < ncomp; c++) for (int j = 0; j < edim; j++) for (int k = 0; k < edim; k++) for (int dof = 0; dof < nd; dof++)
        cv(k + c * edim, dof, i) += T.Ainv[k][j] * rd(dof + c * nd_all, j, i);
    } else if (op == OP_GRAD) {                             // feevaluator_h1.jl:77-94
      for (int i = 0; i < nq; i++) for (int dof = 0; dof < nd; dof++) for (int c = 0; c < ncomp; c++) for (int k = 0; k < edim; k++) {
        for (int j = 0; j < edim; j++) cv(k + c * edim, dof, i) += T.Ainv[k][j] * rd(dof + c * nd_all, j, i);
        cv(k + c * edim, dof, i) *= co(c, dof);
      }
    } else if (op == OP_SYMGRAD) {                          // feevaluator_h1.jl:97-116 (offdiagval = 1)
      const double offdiagval = 1.0;
      for (int i = 0; i < nq; i++) for (int dof = 0; dof < nd; dof++) for (int c = 0; c < ncomp; c++) for (int k = 0; k < edim; k++) for (int j = 0; j < edim; j++) {
        int tgt = compress[k + c * edim] - 1;
        if (k != c) cv(tgt, dof, i) += offdiagval * T.Ainv[k][j] * rd(dof + c * nd_all, j, i);
        else cv(tgt, dof, i) += T.Ainv[k][j] * rd(dof + c * nd_all, j, i);
      }
    } else if (op == OP_DIV && !coeffs_flag) {              // feevaluator_h1.jl:119-130
      for (int i = 0; i < nq; i++) for (int dof = 0; dof < nd; dof++) for (int k = 0; k < edim; k++) for (int j = 0; j < edim; j++)
        cv(0, dof, i) += T.Ainv[k][j] * rd(dof + k * nd_all, j, i);
    } else if (op == OP_DIV) {                              // feevaluator_h1.jl:133-145
      for (int i = 0; i < nq; i++) for (int dof = 0; dof < nd; dof++) for (int k = 0; k < edim; k++) for (int j = 0; j < edim; j++)
        cv(0, dof, i) += T.Ainv[k][j] * rd(dof + k * nd_all, j, i) * co(k, dof);
    }
  }
};

// NeededDerivative4Operator / QuadratureOrderShift4Operator (functionoperators.jl:206-225, 294)
int quadorder_shift(int op) { return (op == OP_GRAD || op == OP_SYMGRAD || op == OP_DIV) ? -1 : 0; }

// -------------------------------------------------------------------------------------
// ExtendableSparseMatrix stand-in (ExtendableSparse.jl semantics): a CSC part searched
// first, then a per-column sorted linked list (LNK); flush! merges LNK into CSC keeping
// explicit zeros.  Indices are 1-based Int64 on the wire.
// -------------------------------------------------------------------------------------
struct ExtSparse {
  i64 m, n;
  std::vector<i64> colptr, rowval; std::vector<double> nzval;     // CSC part (0-based internally)
  std::vector<i64> head; std::vector<i64> lrow, lnext; std::vector<double> lval;  // LNK part
  i64 lnk_nnz = 0;
  ExtSparse(i64 m_, i64 n_) : m(m_), n(n_), colptr(n_ + 1, 0), head(n_, -1) {}
  inline void add(double v, i64 i, i64 j) {                   // rawupdateindex!(A, +, v, i, j), 0-based
    i64 lo = colptr[j], hi = colptr[j + 1];
    if (hi > lo) {
      const i64* p = std::lower_bound(rowval.data() + lo, rowval.data() + hi, i);
      if (p != rowval.data() + hi && *p == i) { nzval[p - rowval.data()] += v; return; }
    }
    i64 k = head[j], k0 = -1;
    while (k >= 0 && lrow[k] < i) { k0 = k; k = lnext[k]; }
    if (k >= 0 && lrow[k] == i) { lval[k] += v; return; }
    i64 nk = (i64)lrow.size();
    lrow.push_back(i); lval.push_back(0.0 + v); lnext.push_back(k);
    if (k0 < 0) head[j] = nk; else lnext[k0] = nk;
    lnk_nnz++;
  }
  void flush() {                                             // flush!: csc = lnk + csc
    if (lnk_nnz == 0) return;
    std::vector<i64> ncp(n + 1, 0), nrow; std::vector<double> nval;
    nrow.reserve(rowval.size() + lnk_nnz); nval.reserve(rowval.size() + lnk_nnz);
    for (i64 j = 0; j < n; j++) {
      i64 a = colptr[j], ae = colptr[j + 1], k = head[j];
      while (a < ae || k >= 0) {
        if (k < 0 || (a < ae && rowval[a] < lrow[k])) { nrow.push_back(rowval[a]); nval.push_back(nzval[a]); a++; }
        else if (a >= ae || lrow[k] < rowval[a]) { nrow.push_back(lrow[k]); nval.push_back(lval[k]); k = lnext[k]; }
        else { nrow.push_back(rowval[a]); nval.push_back(nzval[a] + lval[k]); a++; k = lnext[k]; }
      }
      ncp[j + 1] = (i64)nrow.size();
    }
    colptr.swap(ncp); rowval.swap(nrow); nzval.swap(nval);
    std::fill(head.begin(), head.end(), -1); lrow.clear(); lnext.clear(); lval.clear(); lnk_nnz = 0;
  }
};

// _addnz (src/fematrix.jl:54-58)
inline void addnz(ExtSparse* A, i64 i, i64 j, double v, double fac) {
  if (v != 0.0) A->add(v * fac, i - 1, j - 1);
}
// Magnitude mode (test infrastructure, not part of the reference): the assembly loop accumulates sum |w_q * a_k * b_k| per
// local entry and adds |.| into the matrix, with the zero test of _addnz still made on the true value, so the pattern is
// unchanged.  The result S_ij bounds the rounding error of ANY evaluation order of entry ij by (#ops) * eps * S_ij; the parity
// tests use it to tell entries that are small by cancellation (no order but the reference's own reproduces their digits)
// from entries that are simply small.
static bool g_magnitude_mode = false;
inline void addnz2(ExtSparse* A, i64 i, i64 j, double vtest, double vadd) {
  if (vtest != 0.0) A->add(vadd, i - 1, j - 1);
}

inline bool in_regions(const Grid& g, i64 cell, const i32* regions, int nregions) {
  if (nregions == 1 && regions[0] == 0) return true;        // regions == [0]
  for (int r = 0; r < nregions; r++) if (g.regions && g.regions[cell] == regions[r]) return true;
  return false;
}

void apply_action(int action, const double* p, const double* in, double* out) {
  if (action == ACT_HOOKE2D) {                              // pdeoperators.jl:265-270
    double mu = p[0], la = p[1];
    out[0] = (la + 2 * mu) * in[0] + la * in[1];
    out[1] = (la + 2 * mu) * in[1] + la * in[0];
    out[2] = mu * in[2];
  } else if (action == ACT_HOOKE3D) {                       // pdeoperators.jl:304-312
    double mu = p[0], la = p[1];
    out[0] = (la + 2 * mu) * in[0] + la * (in[1] + in[2]);
    out[1] = (la + 2 * mu) * in[1] + la * (in[0] + in[2]);
    out[2] = (la + 2 * mu) * in[2] + la * (in[0] + in[1]);
    out[3] = mu * in[3]; out[4] = mu * in[4]; out[5] = mu * in[5];
  }
}

int polyorder_of(const Space& s, const Grid& g) { return fe_info(s.fe, s.ncomp, g.dim, g.xdim > g.dim).polyorder; }

}  // namespace

// =====================================================================================
// C ABI (ctypes from oracle/oracle.py)
// =====================================================================================
extern "C" {

const char* orc_last_error() { return g_err.c_str(); }
void orc_set_magnitude_mode(int on) { g_magnitude_mode = on != 0; }

struct orc_grid {
  int dim, xdim; i64 nnodes, ncells, nfaces;
  const double* coords; const i32* cellnodes; const double* cellvolumes; const i32* cellregions;
  const i32* cellfaces; const i32* cellfacesigns; const i32* cellfaceorient; const double* facenormals; const double* facevolumes;
};
struct orc_space { int fetype, ncomp; i64 ndofs; int nd_cell; const i32* celldofs; };

static Grid to_grid(const orc_grid* g) {
  Grid r; r.dim = g->dim; r.xdim = g->xdim; r.nnodes = g->nnodes; r.ncells = g->ncells; r.nfaces = g->nfaces;
  r.coords = g->coords; r.cellnodes = g->cellnodes; r.vol = g->cellvolumes; r.regions = g->cellregions;
  r.cellfaces = g->cellfaces; r.signs = g->cellfacesigns; r.orient = g->cellfaceorient; r.fnormals = g->facenormals; r.fvol = g->facevolumes;
  return r;
}
static Space to_space(const orc_space* s) { Space r; r.fe = s->fetype; r.ncomp = s->ncomp; r.ndofs = s->ndofs; r.nd = s->nd_cell; r.celldofs = s->celldofs; return r; }

// ---- quadrature / reference tables (for cross-checking the host mirror) --------------
void orc_qrule_override(int edim, int order, int nq, const double* xref, const double* w) {
  g_override_edim = edim; g_override_order = order;
  if (nq <= 0) { g_override_edim = g_override_order = -1; return; }
  g_override.dim = edim; g_override.xref.assign(xref, xref + (size_t)nq * edim); g_override.w.assign(w, w + nq);
}
int orc_qrule(int edim, int order, int* nq, double* xref, double* w, int cap) {
  QRule q; if (!make_qrule(edim, order, q)) return -1;
  *nq = q.n();
  if (xref && w) { if (cap < q.n()) { g_err = "capacity"; return -1; } std::memcpy(xref, q.xref.data(), q.xref.size() * 8); std::memcpy(w, q.w.data(), q.w.size() * 8); }
  return 0;
}
// values[i][dof_all][comp], derivs[i][j][dof_all + comp*nd_all] at caller-given reference points
int orc_reftables(int fetype, int ncomp, int edim, int nq, const double* xref, double* values, double* derivs) {
  QRule q; q.dim = edim; q.xref.assign(xref, xref + (size_t)nq * edim); q.w.assign(nq, 0.0);
  FEInfo fi = fe_info(fetype, ncomp, edim); if (fi.nd < 0) { g_err = "unknown FEType"; return -1; }
  for (int i = 0; i < q.n(); i++) {
    RefB<double> rb(fi.nd_all, fi.ncomp);
    eval_basis<double>(fetype, rb, &q.xref[(size_t)i * edim], edim, ncomp);
    for (int dof = 0; dof < fi.nd_all; dof++) for (int c = 0; c < fi.ncomp; c++) values[((size_t)i * fi.nd_all + dof) * fi.ncomp + c] = rb(dof, c);
    Dual x[3];
    for (int j = 0; j < edim; j++) { x[j] = Dual(q.xref[(size_t)i * edim + j]); x[j].d[j] = 1.0; }
    RefB<Dual> rbd(fi.nd_all, fi.ncomp);
    eval_basis<Dual>(fetype, rbd, x, edim, ncomp);
    for (int c = 0; c < fi.ncomp; c++) for (int dof = 0; dof < fi.nd_all; dof++) for (int j = 0; j < edim; j++)
      derivs[((size_t)i * edim + j) * (fi.nd_all * fi.ncomp) + dof + c * fi.nd_all] = rbd(dof, c).d[j];
  }
  return 0;
}

// ---- matrix handle -------------------------------------------------------------------
void* orc_matrix_create(i64 m, i64 n) { return new ExtSparse(m, n); }
void orc_matrix_destroy(void* A) { delete (ExtSparse*)A; }
void orc_matrix_flush(void* A) { ((ExtSparse*)A)->flush(); }
i64 orc_matrix_nnz(void* A) { ExtSparse* a = (ExtSparse*)A; return (i64)a->rowval.size(); }
// fill!(A, 0) keeps the pattern (src/fematrix.jl:220-232)
void orc_matrix_fill_zero(void* A) { ExtSparse* a = (ExtSparse*)A; std::fill(a->nzval.begin(), a->nzval.end(), 0.0); }
// 1-based Int64 colptr/rowval like SparseMatrixCSC{Float64,Int64}
void orc_matrix_get(void* A, i64* colptr, i64* rowval, double* nzval) {
  ExtSparse* a = (ExtSparse*)A;
  for (i64 j = 0; j <= a->n; j++) colptr[j] = a->colptr[j] + 1;
  for (size_t k = 0; k < a->rowval.size(); k++) { rowval[k] = a->rowval[k] + 1; nzval[k] = a->nzval[k]; }
}

// ---- fixed (coefficient) argument of a trilinear form: assemble!(A, AP, FEB; fixed_arguments = [1]) with nFE = 3
// (bilinearform.jl:235-257: the operator evaluation of FEB[1] at every quadrature point is the first part of the action input).
// Set before orc_blf_assemble, cleared with sa = NULL.  Used with ACT_CONVECTION, the kernel of ConvectionOperator(a_from, a_operator,
// xdim, ncomponents; a_to = 1) (pdeoperators.jl:435-510): result[j] = sum_k input[k] * input[xdim + (j-1) xdim + k].
static orc_space g_fixed_space; static bool g_fixed_on = false; static int g_fixed_op = 0; static const double* g_fixed_coeffs = nullptr;
void orc_set_fixed_argument(const orc_space* sa, int op_a, const double* coeffs) {
  g_fixed_on = sa != nullptr;
  if (sa) { g_fixed_space = *sa; g_fixed_op = op_a; g_fixed_coeffs = coeffs; }
}

// ---- BilinearForm assemble! (src/assemblypatterns/bilinearform.jl:92-380) ------------
// apply_action_to == [1] (all operators on the path, pdeoperators.jl:164,188,222,272)
int orc_blf_assemble(void* Aptr, const orc_grid* og, const orc_space* os1, const orc_space* os2, int op1, int op2,
                     int action, const double* act_params, int apt, const i32* regions, int nregions,
                     double factor, int transposed_assembly, void* transpose_copy, double factor_transpose,
                     i64 offsetX, i64 offsetY, int bonus_quadorder) {
  ExtSparse* A = (ExtSparse*)Aptr; ExtSparse* At = (ExtSparse*)transpose_copy;
  Grid g = to_grid(og); Space s1 = to_space(os1), s2 = to_space(os2);
  int edim = g.dim;
  // prepare_assembly! : quadrature order (assemblypatterns.jl:559-565)
  int quadorder = bonus_quadorder + polyorder_of(s1, g) + quadorder_shift(op1) + polyorder_of(s2, g) + quadorder_shift(op2);
  Space sa{}; Evaluator ea;
  if (g_fixed_on) { sa = to_space(&g_fixed_space); quadorder += polyorder_of(sa, g) + quadorder_shift(g_fixed_op); }   // all FE of the pattern, 559-565
  if (quadorder < 0) quadorder = 0;
  QRule q; if (!make_qrule(edim, quadorder, q)) return -1;
  Evaluator e1, e2store; Evaluator* e2 = &e2store;
  if (!e1.init(&g, s1, op1, q)) return -1;
  if (g_fixed_on && !ea.init(&g, sa, g_fixed_op, q)) return -1;
  if ((action == ACT_CONVECTION) != g_fixed_on) { g_err = "ACT_CONVECTION needs a fixed argument (and vice versa)"; return -1; }
  bool same = (s1.celldofs == s2.celldofs && s1.fe == s2.fe && s1.ncomp == s2.ncomp && op1 == op2);  // evaluator reuse 567-584
  if (same) e2 = &e1; else if (!e2store.init(&g, s2, op2, q)) return -1;
  int nd1 = e1.nd, nd2 = e2->nd, nq = q.n();
  int rdim_action;   // action_resultdim
  int in_dim = e1.resultdim;
  const int adim = g_fixed_on ? ea.resultdim : 0;      // offsets[nFE-1] = basisdim of the fixed argument (150-152)
  if (action == ACT_NONE) rdim_action = e1.resultdim;
  else if (action == ACT_CONVECTION) {
    if (adim < 1 || in_dim % adim) { g_err = "convection: ansatz operator length is not a multiple of the coefficient length"; return -1; }
    rdim_action = in_dim / adim;
  }
  else { rdim_action = (action == ACT_HOOKE2D) ? 3 : 6; if (in_dim != rdim_action) { g_err = "action/operator size mismatch"; return -1; } }
  std::vector<double> aq((size_t)std::max(adim, 1) * q.n(), 0.0), ca(g_fixed_on ? ea.nd : 0);
  if (g_fixed_on && g_magnitude_mode) { g_err = "magnitude mode: no fixed arguments"; return -1; }
  if (rdim_action != e2->resultdim) { g_err = "operator result dimensions do not match"; return -1; }
  std::vector<double> local((size_t)nd1 * nd2, 0.0), action_result(rdim_action), action_input(in_dim);
  const bool mag = g_magnitude_mode;
  std::vector<double> localabs(mag ? (size_t)nd1 * nd2 : 0, 0.0), action_abs(rdim_action), input_abs(in_dim), params_abs(2, 0.0);
  if (mag && act_params) { params_abs[0] = std::fabs(act_params[0]); params_abs[1] = std::fabs(act_params[1]); }
  if (mag && transpose_copy) { g_err = "magnitude mode: no transpose copy"; return -1; }
  const bool is_symmetric = (apt == APT_SYMMETRIC);
  for (i64 item = 0; item < g.ncells; item++) {
    if (!in_regions(g, item, regions, nregions)) continue;
    e1.update(item); if (e2 != &e1) e2->update(item);       // update_assembly! (assemblypatterns.jl:251-276)
    if (g_fixed_on) {                                       // 235-257: FEB[1] evaluated at the quadrature points, eval_febe! from 0 in dof order
      ea.update(item);
      const i32* da = sa.celldofs + item * ea.nd;
      for (int d = 0; d < ea.nd; d++) ca[d] = g_fixed_coeffs[da[d] - 1] * 1.0;
      std::fill(aq.begin(), aq.end(), 0.0);
      for (int i = 0; i < nq; i++)
        for (int d = 0; d < ea.nd; d++)
          for (int k = 0; k < adim; k++) aq[(size_t)i * adim + k] += ca[d] * ea.cv(k, d, i) * 1;
    }
    const bool locsym = is_symmetric;                       // dofitems[1] == dofitems[2] for continuous operators
    for (int i = 0; i < nq; i++) {
      for (int di = 0; di < nd1; di++) {
        if (action == ACT_NONE) {                           // 306-308
          for (int k = 0; k < rdim_action; k++) action_result[k] = e1.cv(k, di, i) * 1.0 * 1.0;
        } else if (action == ACT_CONVECTION) {              // 310-313 with the input [a(x_i), operator evaluation of dof di]
          for (int k = 0; k < in_dim; k++) action_input[k] = e1.cv(k, di, i) * 1.0 * 1.0;
          const double* a = &aq[(size_t)i * adim];
          for (int j = 0; j < rdim_action; j++) {           // pdeoperators.jl:459-467
            double r = 0;
            for (int k = 0; k < adim; k++) r += a[k] * action_input[(size_t)j * adim + k];
            action_result[j] = r;
          }
        } else {                                            // 310-313
          for (int k = 0; k < in_dim; k++) action_input[k] = e1.cv(k, di, i) * 1.0 * 1.0;
          apply_action(action, act_params, action_input.data(), action_result.data());
        }
        if (mag) {
          if (action == ACT_NONE) for (int k = 0; k < rdim_action; k++) action_abs[k] = std::fabs(action_result[k]);
          else { for (int k = 0; k < in_dim; k++) input_abs[k] = std::fabs(action_input[k]); apply_action(action, params_abs.data(), input_abs.data(), action_abs.data()); }
          for (int dj = (apt == APT_LUMPED ? di : (locsym ? di : 0)); dj < (apt == APT_LUMPED ? di + 1 : nd2); dj++) {
            double t = 0; for (int k = 0; k < rdim_action; k++) t += action_abs[k] * std::fabs(e2->cv(k, dj, i));
            localabs[(size_t)di * nd2 + dj] += std::fabs(q.w[i]) * t;
          }
        }
        // basismul! (180-220)
        if (apt == APT_LUMPED) {
          double t = 0; for (int k = 0; k < rdim_action; k++) t += action_result[k] * e2->cv(k, di, i);
          local[(size_t)di * nd2 + di] += q.w[i] * t;
        } else {
          for (int dj = (locsym ? di : 0); dj < nd2; dj++) {
            double t = 0; for (int k = 0; k < rdim_action; k++) t += action_result[k] * e2->cv(k, dj, i);
            local[(size_t)di * nd2 + dj] += q.w[i] * t;
          }
        }
      }
    }
    double itemfactor = g.vol[item] * factor * 1.0;          // 320
    const i32* d1 = s1.celldofs + item * nd1; const i32* d2 = s2.celldofs + item * nd2;
    if (mag) {
      const double af = std::fabs(itemfactor);
      for (int di = 0; di < nd1; di++) for (int dj = 0; dj < nd2; dj++) {
        if (locsym && dj < di) continue;
        if (apt == APT_LUMPED && dj != di) continue;
        const double v = local[(size_t)di * nd2 + dj] * itemfactor, va = localabs[(size_t)di * nd2 + dj] * af;
        i64 arow = d1[di] + offsetX, acol = d2[dj] + offsetY;
        if (!locsym && transposed_assembly) std::swap(arow, acol);
        addnz2(A, arow, acol, v, va);
        if (locsym && dj != di) addnz2(A, d1[dj] + offsetX, d2[di] + offsetY, v, va);
      }
      std::fill(local.begin(), local.end(), 0.0);
      std::fill(localabs.begin(), localabs.end(), 0.0);
      continue;
    }
    if (locsym) {                                           // 329-346
      for (int di = 0; di < nd1; di++) for (int dj = di + 1; dj < nd2; dj++) {
        double v = local[(size_t)di * nd2 + dj] * itemfactor;
        addnz(A, d1[di] + offsetX, d2[dj] + offsetY, v, 1);
        addnz(A, d1[dj] + offsetX, d2[di] + offsetY, v, 1);
      }
      for (int di = 0; di < nd1; di++) addnz(A, d1[di] + offsetX, d2[di] + offsetY, local[(size_t)di * nd2 + di] * itemfactor, 1);
    } else {                                                // 347-367 (the lumped branch test at 323 is never true)
      for (int di = 0; di < nd1; di++) {
        i64 arow = d1[di] + offsetX;
        for (int dj = 0; dj < nd2; dj++) {
          i64 acol = d2[dj] + offsetY;
          double v = local[(size_t)di * nd2 + dj] * itemfactor;
          if (transposed_assembly) addnz(A, acol, arow, v, 1); else addnz(A, arow, acol, v, 1);
          if (At) {
            double vt = local[(size_t)di * nd2 + dj] * itemfactor / factor * factor_transpose;
            if (transposed_assembly) addnz(At, arow, acol, vt, -1); else addnz(At, acol, arow, vt, -1);
          }
        }
      }
    }
    std::fill(local.begin(), local.end(), 0.0);             // 369
  }
  return 0;
}

// ---- NonlinearForm full_assemble! (src/assemblypatterns/nonlinearform.jl:44-245) for the Newton form of the convection term:
// ConvectionOperator(a_from, a_operator, xdim, ncomponents; newton = true) = NonlinearForm(test_operator, [a_operator, ansatz_operator],
// [a_from, a_from], convection_function_fe_1, argsizes; jacobian = convection_jacobian) (pdeoperators.jl:459-493): all three FESpaces are the
// space of the unknown u, operators [op_a (Identity), op_g (Gradient), op_t (test)], newton_args = [1, 2].
//   input_i = [op_a(u)(x_i), op_g(u)(x_i)] (eval_febe! from 0 in dof order, 124-141);
//   value[j] = sum_k input[k] input[xdim + (j-1) xdim + k];  jac[j,k] = input[xdim + (j-1) xdim + k], jac[j, xdim + (j-1) xdim + k] = input[k]
//   (sparse jacobian: mul! runs over the stored columns in ascending order, plain multiply and add);
//   matrix: for every ansatz dof: input2 = [op_a(phi_dof), op_g(phi_dof)], result = jac input2, local[dof_i,dof_j] += (result . op_t(v_dof_j)) w_i (167-187);
//   rhs:    result = jac input_i, localb[dof_j] += ((result - value) . op_t(v_dof_j)) w_i (189-201);
//   _addnz(A, acol, arow, local, itemfactor) -- the zero test is on the UNSCALED local entry (216-221); localb .*= itemfactor; b += localb (226-233).
int orc_nlf_convection(void* Aptr, double* b, const orc_grid* og, const orc_space* os, int op_a, int op_g, int op_t, const double* coeffs,
                       const i32* regions, int nregions, double factor, int transposed_assembly, i64 offsetX, i64 offsetY, int bonus_quadorder) {
  ExtSparse* A = (ExtSparse*)Aptr;
  Grid g = to_grid(og); Space s = to_space(os);
  const int edim = g.dim;
  int quadorder = bonus_quadorder + 3 * polyorder_of(s, g) + quadorder_shift(op_a) + quadorder_shift(op_g) + quadorder_shift(op_t);
  if (quadorder < 0) quadorder = 0;
  QRule q; if (!make_qrule(edim, quadorder, q)) return -1;
  Evaluator ea, eg, etst; Evaluator* et = &etst;
  if (!ea.init(&g, s, op_a, q) || !eg.init(&g, s, op_g, q)) return -1;
  if (op_t == op_a) et = &ea; else if (op_t == op_g) et = &eg; else if (!etst.init(&g, s, op_t, q)) return -1;   // evaluator reuse (assemblypatterns.jl:567-584)
  const int nd = ea.nd, nq = q.n(), xdim = ea.resultdim, gdim = eg.resultdim;
  if (xdim < 1 || gdim % xdim) { g_err = "convection: operator lengths do not fit"; return -1; }
  const int nc = gdim / xdim;
  if (et->resultdim != nc) { g_err = "convection: test operator length"; return -1; }
  const int nin = xdim + gdim;
  std::vector<double> input((size_t)nq * nin), in2(nin), res(nc), value(nc), local((size_t)nd * nd, 0.0), localb(nd, 0.0), c(nd);
  auto jacmul = [&](const double* in, const double* x, double* y) {     // y = jac(in) x, stored columns ascending
    for (int j = 0; j < nc; j++) y[j] = 0.0;
    for (int k = 0; k < xdim; k++)
      for (int j = 0; j < nc; j++) y[j] += in[xdim + j * xdim + k] * x[k];
    for (int j = 0; j < nc; j++)
      for (int k = 0; k < xdim; k++) y[j] += in[k] * x[xdim + j * xdim + k];
  };
  for (i64 item = 0; item < g.ncells; item++) {
    if (!in_regions(g, item, regions, nregions)) continue;
    ea.update(item); eg.update(item); if (et != &ea && et != &eg) et->update(item);
    const i32* dofs = s.celldofs + item * nd;
    for (int d = 0; d < nd; d++) c[d] = coeffs[dofs[d] - 1] * 1.0;
    std::fill(input.begin(), input.end(), 0.0);
    for (int i = 0; i < nq; i++) {
      double* in = &input[(size_t)i * nin];
      for (int d = 0; d < nd; d++) for (int k = 0; k < xdim; k++) in[k] += c[d] * ea.cv(k, d, i) * 1;
      for (int d = 0; d < nd; d++) for (int k = 0; k < gdim; k++) in[xdim + k] += c[d] * eg.cv(k, d, i) * 1;
    }
    for (int i = 0; i < nq; i++) {
      const double* in = &input[(size_t)i * nin];
      for (int j = 0; j < nc; j++) { double r = 0; for (int k = 0; k < xdim; k++) r += in[k] * in[xdim + j * xdim + k]; value[j] = r; }
      for (int di = 0; di < nd; di++) {
        for (int k = 0; k < xdim; k++) in2[k] = ea.cv(k, di, i) * 1;
        for (int k = 0; k < gdim; k++) in2[xdim + k] = eg.cv(k, di, i) * 1;
        jacmul(in, in2.data(), res.data());
        for (int dj = 0; dj < nd; dj++) {
          double t = 0; for (int k = 0; k < nc; k++) t += res[k] * et->cv(k, dj, i);
          local[(size_t)di * nd + dj] += t * q.w[i];
        }
      }
      for (int dj = 0; dj < nd; dj++) {
        jacmul(in, in, res.data());
        double t = 0; for (int k = 0; k < nc; k++) t += (res[k] - value[k]) * et->cv(k, dj, i);
        localb[dj] += t * q.w[i];
      }
    }
    const double itemfactor = g.vol[item] * factor * 1.0;
    for (int di = 0; di < nd; di++) {
      const i64 arow = dofs[di] + offsetY;
      for (int dj = 0; dj < nd; dj++) {
        const i64 acol = dofs[dj] + offsetX;
        const double v = local[(size_t)di * nd + dj];
        if (transposed_assembly) addnz(A, acol, arow, v, itemfactor); else addnz(A, arow, acol, v, itemfactor);
      }
    }
    std::fill(local.begin(), local.end(), 0.0);
    if (b) for (int dj = 0; dj < nd; dj++) { localb[dj] *= itemfactor; b[dofs[dj] - 1 + offsetX] += localb[dj]; }
    std::fill(localb.begin(), localb.end(), 0.0);
  }
  return 0;
}

// ---- LinearForm assemble! (src/assemblypatterns/linearform.jl:47-237), nFE == 1 ------
// fsrc: F_NONE -> no action (action_input = ones, 74-75); F_CONST -> fdata[resultdim];
// F_QP_TABLE -> fdata[cell][qp][resultdim] (fdot_action evaluated by the host, actions.jl:119-128)
int orc_lf_assemble(double* b, const orc_grid* og, const orc_space* os, int op, int fsrc, const double* fdata,
                    const i32* regions, int nregions, double factor, i64 offset, int bonus_quadorder, int* nq_out) {
  Grid g = to_grid(og); Space s = to_space(os);
  int edim = g.dim;
  int quadorder = bonus_quadorder + polyorder_of(s, g) + quadorder_shift(op);
  if (quadorder < 0) quadorder = 0;
  QRule q; if (!make_qrule(edim, quadorder, q)) return -1;
  if (nq_out) *nq_out = q.n();
  if (!b) return 0;
  Evaluator e; if (!e.init(&g, s, op, q)) return -1;
  int nd = e.nd, nq = q.n(), rdim = e.resultdim;
  std::vector<double> localb(nd, 0.0), ones(rdim, 1.0);
  for (i64 item = 0; item < g.ncells; item++) {
    if (!in_regions(g, item, regions, nregions)) continue;
    e.update(item);
    for (int i = 0; i < nq; i++) {
      const double* r = (fsrc == F_NONE) ? ones.data() : (fsrc == F_CONST) ? fdata : fdata + ((size_t)item * nq + i) * rdim;
      for (int d = 0; d < nd; d++) {                        // 181-210
        double t = 0; for (int k = 0; k < rdim; k++) t += r[k] * e.cv(k, d, i);
        localb[d] += t * q.w[i];
      }
    }
    double itemfactor = factor * g.vol[item] * 1.0;          // 215
    const i32* dofs = s.celldofs + item * nd;
    for (int d = 0; d < nd; d++) b[dofs[d] - 1 + offset] += localb[d] * itemfactor;   // 216-220
    std::fill(localb.begin(), localb.end(), 0.0);
  }
  return 0;
}

// ---- ItemIntegrator evaluate! / evaluate (src/assemblypatterns/itemintegrator.jl:160-300, 316-360), one argument ------------
// kind: II_NONE  -> NoAction: b[j,item] += input_i[j] * w_i * |T|                                   (262-268)
//       II_L2NORM -> L2NormIntegrator kernel (99-108): temp[j] = 0 + input[j]; result = sum_j temp[j]^2
//       II_L2ERROR-> L2ErrorIntegrator kernel (52-69): val[j] = data(x_i)[j] - input[j]*factor; result = sum_j val[j]^2
//                   (data[cell][qp][ncomp] = compare_data evaluated by the host at the quadrature points)
// input_i[k] = sum_dof coeffs[dof] * cvals[k,dof,i] * 1 accumulated from 0 in dof order (eval_febe!, feevaluator.jl:445-452).
// b may be NULL; total[resultdim] (may be NULL) is what evaluate() returns: every (item, qp) term added to ONE running sum in
// loop order (AccumulatingVector, 346-352).
int orc_ii_evaluate(double* b, double* total, const orc_grid* og, const orc_space* os, int op, int kind, const double* coeffs,
                    double factor, const double* data, const i32* regions, int nregions, int bonus_quadorder, int* nq_out,
                    int* resultdim_out) {
  Grid g = to_grid(og); Space s = to_space(os);
  int edim = g.dim;
  int quadorder = bonus_quadorder + polyorder_of(s, g) + quadorder_shift(op);
  if (quadorder < 0) quadorder = 0;
  QRule q; if (!make_qrule(edim, quadorder, q)) return -1;
  if (nq_out) *nq_out = q.n();
  Evaluator e; if (!e.init(&g, s, op, q)) return -1;
  const int nd = e.nd, nq = q.n(), rdim = e.resultdim;
  const int ardim = (kind == II_NONE) ? rdim : 1;
  if (resultdim_out) *resultdim_out = ardim;
  if (!coeffs) return 0;
  std::vector<double> input(rdim, 0.0), res(ardim, 0.0), c(nd, 0.0);
  if (total) for (int j = 0; j < ardim; j++) total[j] = 0.0;
  for (i64 item = 0; item < g.ncells; item++) {
    if (!in_regions(g, item, regions, nregions)) continue;
    e.update(item);
    const i32* dofs = s.celldofs + item * nd;
    for (int d = 0; d < nd; d++) c[d] = coeffs[dofs[d] - 1] * 1.0;     // get_coeffs! ; coeffs .*= coeff4dofitem (234-236)
    for (int i = 0; i < nq; i++) {
      std::fill(input.begin(), input.end(), 0.0);
      for (int d = 0; d < nd; d++)
        for (int k = 0; k < rdim; k++) input[k] += c[d] * e.cv(k, d, i) * 1;
      if (kind == II_NONE) {
        for (int j = 0; j < rdim; j++) res[j] = input[j];
      } else if (kind == II_L2NORM) {
        double r = 0;
        for (int j = 0; j < rdim; j++) { double t = 0.0; t += input[j]; r += t * t; }
        res[0] = r;
      } else {
        const double* dv = data + ((size_t)item * nq + i) * rdim;
        double r = 0;
        for (int j = 0; j < rdim; j++) { double v = dv[j]; v -= input[j] * factor; r += v * v; }
        res[0] = r;
      }
      for (int j = 0; j < ardim; j++) {
        const double term = res[j] * q.w[i] * g.vol[item];               // 266 / 291
        if (b) b[(size_t)item * ardim + j] += term;
        if (total) total[j] += term;
      }
    }
  }
  return 0;
}

// operator evaluation of an FE function at the quadrature points of every cell (eval_febe!, feevaluator.jl:445-452, as the assembly
// loops evaluate their coefficient arguments: linearform.jl:141-160, bilinearform.jl:235-257): table[cell][qp][resultdim]
int orc_feb_table(const orc_grid* og, const orc_space* os, int op, const double* coeffs, int order, double* table, int* nq_out, int* rd_out) {
  Grid g = to_grid(og); Space s = to_space(os);
  QRule q; if (!make_qrule(g.dim, order < 0 ? 0 : order, q)) return -1;
  Evaluator e; if (!e.init(&g, s, op, q)) return -1;
  if (nq_out) *nq_out = q.n();
  if (rd_out) *rd_out = e.resultdim;
  if (!table) return 0;
  const int nd = e.nd, nq = q.n(), rd = e.resultdim;
  std::vector<double> c(nd);
  for (i64 item = 0; item < g.ncells; item++) {
    e.update(item);
    const i32* dofs = s.celldofs + item * nd;
    for (int d = 0; d < nd; d++) c[d] = coeffs[dofs[d] - 1] * 1.0;
    for (int i = 0; i < nq; i++) {
      double* t = table + ((size_t)item * nq + i) * rd;
      for (int k = 0; k < rd; k++) t[k] = 0.0;
      for (int d = 0; d < nd; d++)
        for (int k = 0; k < rd; k++) t[k] += c[d] * e.cv(k, d, i) * 1;
    }
  }
  return 0;
}

// physical quadrature points x = b + A*xref of every cell (eval_trafo!, linearform.jl:197-201):
// xq[cell][qp][dim]; used by tests to tabulate f independently of the host mirror
int orc_quadpoints(const orc_grid* og, int order, double* xq) {
  Grid g = to_grid(og); QRule q; if (!make_qrule(g.dim, order, q)) return -1;
  Trafo T; int d = g.dim;
  for (i64 c = 0; c < g.ncells; c++) {
    T.update(g, c);
    for (int i = 0; i < q.n(); i++) for (int k = 0; k < d; k++) {
      double x = T.b[k];
      for (int j = 0; j < d; j++) x += T.A[k][j] * q.xref[(size_t)i * d + j];
      xq[((size_t)c * q.n() + i) * d + k] = x;
    }
  }
  return 0;
}

}  // extern "C"
