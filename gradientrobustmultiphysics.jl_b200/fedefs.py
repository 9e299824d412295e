"""Finite element definitions of the hot path (host mirror of src/fedefs/*.jl).

Each FEType carries what the reference's definition files provide:
  get_ncomponents, get_ndofs, get_ndofs_all, get_polynomialorder, get_dofmap_pattern,
  get_basis (the reference-cell closure), and flags telling the evaluator whether
  per-cell coefficients / subsets apply (AbstractH1FiniteElementWithCoefficients,
  AbstractHdivFiniteElement).

The basis closures are written over a generic scalar so that they can be evaluated with
floats (reference values, feevaluator.jl:64-68) and with forward-mode dual numbers
(reference jacobians, feevaluator.jl:235-293 -- ForwardDiff in the reference; `Dual`
below restates its un-fused product rule).  The tables produced here are *inputs* of
libgrmp_cuda (in production they come straight out of the Julia FEEvaluator).
"""
from __future__ import annotations

import numpy as np

# integer codes shared with include/grmp.h
FE_H1P1, FE_H1P2, FE_H1BR, FE_HDIVRT0, FE_HDIVBDM1, FE_L2P0 = 1, 2, 3, 4, 5, 6


class Dual:
    """value + up to 3 partial derivatives; (a*b)' = b.v*a' + a.v*b' with separate mul/add."""
    __slots__ = ("v", "d")

    def __init__(self, v, d=None):
        self.v = float(v)
        self.d = np.zeros(3) if d is None else d

    @staticmethod
    def _lift(x):
        return x if isinstance(x, Dual) else Dual(x)

    def __add__(self, o):
        o = Dual._lift(o)
        return Dual(self.v + o.v, self.d + o.d)

    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, Dual):
            return Dual(self.v - o.v, self.d - o.d)
        return Dual(self.v - o, self.d.copy())

    def __rsub__(self, o):
        return Dual(o - self.v, -self.d)

    def __neg__(self):
        return Dual(-self.v, -self.d)

    def __mul__(self, o):
        if isinstance(o, Dual):
            return Dual(self.v * o.v, (o.v * self.d) + (self.v * o.d))
        return Dual(self.v * o, self.d * o)

    def __rmul__(self, o):
        return Dual(o * self.v, o * self.d)

    def __truediv__(self, o):
        return Dual(self.v / o, self.d / o)


class _RefB:
    """refbasis[dof, comp] (0-based) with Julia's `refbasis[end]` scratch semantics."""

    def __init__(self, nd, nc, zero):
        self.a = [[zero for _ in range(nc)] for _ in range(nd)]
        self.nd, self.nc = nd, nc

    def __getitem__(self, ij):
        return self.a[ij[0]][ij[1]]

    def __setitem__(self, ij, v):
        self.a[ij[0]][ij[1]] = v

    @property
    def last(self):
        return self.a[self.nd - 1][self.nc - 1]

    @last.setter
    def last(self, v):
        self.a[self.nd - 1][self.nc - 1] = v


def _nn(edim):
    return edim + 1


def _ne(edim):
    return {1: 1, 2: 3, 3: 6}[edim]      # Edge1D: the interior dof takes the edge slot


class FEType:
    code = 0
    coefficients = False      # AbstractH1FiniteElementWithCoefficients / Hdiv
    hdiv = False
    broken = False
    name = "?"

    def ndofs(self, edim):            # get_ndofs(ON_CELLS, FEType, EG)
        raise NotImplementedError

    def ndofs_all(self, edim):        # get_ndofs_all
        return self.ndofs(edim)

    def __repr__(self):
        return self.name

    def __eq__(self, o):
        return type(self) is type(o) and self.__dict__ == o.__dict__

    def __hash__(self):
        return hash((type(self).__name__, tuple(sorted(self.__dict__.items()))))


class H1P1(FEType):
    """src/fedefs/h1_p1.jl"""
    code = FE_H1P1

    def __init__(self, ncomponents=1):
        self.ncomponents = ncomponents
        self.name = f"H1P1{{{ncomponents}}}"

    def ndofs(self, edim):
        return _nn(edim) * self.ncomponents

    def polynomialorder(self, edim):
        return 1

    def dofmap_pattern(self, edim):
        return "N1"

    def bface_dofmap_pattern(self, fdim):   # h1_p1.jl:29
        return "N1"

    def basis(self, rb, x, edim):     # h1_p1.jl:64-75
        for k in range(1, self.ncomponents + 1):
            r = (edim + 1) * k - edim - 1
            rb[r, k - 1] = Dual(1.0) if isinstance(x[0], Dual) else 1.0
            for j in range(1, edim + 1):
                rb[r, k - 1] = rb[r, k - 1] - x[j - 1]
                rb[r + j, k - 1] = x[j - 1]


class H1P2(FEType):
    """src/fedefs/h1_p2.jl (also serves H1Pk{n,2,2}: identical floating-point tables)"""
    code = FE_H1P2

    def __init__(self, ncomponents=1, edim=2):
        self.ncomponents, self.edim = ncomponents, edim
        self.name = f"H1P2{{{ncomponents},{edim}}}"

    def ndofs(self, edim):
        return (_nn(edim) + _ne(edim)) * self.ncomponents

    def polynomialorder(self, edim):
        return 2

    def dofmap_pattern(self, edim):   # h1_p2.jl:107-113
        return "N1F1" if edim == 2 else "N1E1"

    def bface_dofmap_pattern(self, fdim):   # h1_p2.jl:37-38
        return "N1I1" if fdim == 1 else "N1E1"

    def basis(self, rb, x, edim):     # h1_p2.jl:123-132 (Edge1D), 208-239
        if edim == 1:
            rb.last = 1.0 - x[0]
            for k in range(1, self.ncomponents + 1):
                l = rb.last
                rb[3 * k - 3, k - 1] = 2.0 * l * (l - 0.5)
                rb[3 * k - 2, k - 1] = 2.0 * x[0] * (x[0] - 0.5)
                rb[3 * k - 1, k - 1] = 4.0 * l * x[0]
        elif edim == 2:
            rb.last = 1.0 - x[0] - x[1]
            for k in range(1, self.ncomponents + 1):
                l = rb.last
                rb[6 * k - 6, k - 1] = 2.0 * l * (l - 0.5)
                rb[6 * k - 5, k - 1] = 2.0 * x[0] * (x[0] - 0.5)
                rb[6 * k - 4, k - 1] = 2.0 * x[1] * (x[1] - 0.5)
                rb[6 * k - 3, k - 1] = 4.0 * l * x[0]
                rb[6 * k - 2, k - 1] = 4.0 * x[0] * x[1]
                rb[6 * k - 1, k - 1] = 4.0 * x[1] * l
        else:
            rb.last = 1.0 - x[0] - x[1] - x[2]
            for k in range(1, self.ncomponents + 1):
                l = rb.last
                rb[10 * k - 10, k - 1] = 2.0 * l * (l - 0.5)
                rb[10 * k - 9, k - 1] = 2.0 * x[0] * (x[0] - 0.5)
                rb[10 * k - 8, k - 1] = 2.0 * x[1] * (x[1] - 0.5)
                rb[10 * k - 7, k - 1] = 2.0 * x[2] * (x[2] - 0.5)
                rb[10 * k - 6, k - 1] = 4.0 * l * x[0]
                rb[10 * k - 5, k - 1] = 4.0 * l * x[1]
                rb[10 * k - 4, k - 1] = 4.0 * l * x[2]
                rb[10 * k - 3, k - 1] = 4.0 * x[0] * x[1]
                rb[10 * k - 2, k - 1] = 4.0 * x[0] * x[2]
                rb[10 * k - 1, k - 1] = 4.0 * x[1] * x[2]


def H1Pk(ncomponents, edim, order):
    """H1Pk{n,e,order} (src/fedefs/h1_pk.jl): orders 1 and 2 coincide with H1P1/H1P2
    (same dof pattern "N1"/"N1F1", and the order-2 closure 171-269 yields bit-identical
    values and ForwardDiff partials: it differs from H1P2 only by power-of-two scalings)."""
    if order == 1:
        return H1P1(ncomponents)
    if order == 2 and edim == 2:
        return H1P2(ncomponents, edim)
    raise NotImplementedError("H1Pk with order >= 3 needs sign-dependent subsets (out of scope, SURVEY.md 2)")


class H1BR(FEType):
    """Bernardi--Raugel, src/fedefs/h1v_br.jl"""
    code = FE_H1BR
    coefficients = True

    def __init__(self, edim=2):
        self.edim = edim
        self.ncomponents = edim
        self.name = f"H1BR{{{edim}}}"

    def ndofs(self, edim):
        return _nn(edim) + _nn(edim) * edim

    def polynomialorder(self, edim):
        return 2 if edim == 2 else 3

    def dofmap_pattern(self, edim):
        return "N1f1"

    def basis(self, rb, x, edim):     # h1v_br.jl:117-130, 218-232
        H1P1(edim).basis(rb, x, edim)
        if edim == 2:
            o = 6
            rb[o + 0, 0] = 6.0 * x[0] * rb[0, 0]
            rb[o + 1, 0] = 6.0 * x[1] * x[0]
            rb[o + 2, 0] = 6.0 * rb[0, 0] * x[1]
            for j in range(3):
                rb[o + j, 1] = rb[o + j, 0]
        else:
            o = 12
            rb[o + 0, 0] = 60.0 * x[0] * rb[0, 0] * x[1]
            rb[o + 1, 0] = 60.0 * rb[0, 0] * x[0] * x[2]
            rb[o + 2, 0] = 60.0 * x[0] * x[1] * x[2]
            rb[o + 3, 0] = 60.0 * rb[0, 0] * x[1] * x[2]
            for j in range(4):
                for k in (1, 2):
                    rb[o + j, k] = rb[o + j, 0]


class HDIVRT0(FEType):
    """src/fedefs/hdiv_rt0.jl"""
    code = FE_HDIVRT0
    coefficients = True
    hdiv = True

    def __init__(self, edim=2):
        self.edim = edim
        self.ncomponents = edim
        self.name = f"HDIVRT0{{{edim}}}"

    def ndofs(self, edim):            # hdiv_rt0.jl:21-22
        return 1 if edim < self.edim else _nn(edim)

    def ncomponents_on(self, edim):   # the face basis is the (scalar) normal flux
        return 1 if edim < self.edim else self.ncomponents

    def polynomialorder(self, edim):  # hdiv_rt0.jl:24-27
        return 0 if edim < self.edim else 1

    def dofmap_pattern(self, edim):
        return "f1"

    def bface_dofmap_pattern(self, fdim):   # hdiv_rt0.jl:30
        return "i1"

    def basis(self, rb, x, edim):     # hdiv_rt0.jl:61-65 (faces), 67-73, 84-92
        if edim < self.edim:
            rb[0, 0] = 1.0
        elif edim == 2:
            rb[0, 0] = x[0];        rb[0, 1] = x[1] - 1.0
            rb[1, 0] = x[0];        rb[1, 1] = x[1]
            rb[2, 0] = x[0] - 1.0;  rb[2, 1] = x[1]
        else:
            rb[0, 0] = 2.0 * x[0];          rb[0, 1] = 2.0 * x[1];          rb[0, 2] = 2.0 * (x[2] - 1.0)
            rb[1, 0] = 2.0 * x[0];          rb[1, 1] = 2.0 * (x[1] - 1.0);  rb[1, 2] = 2.0 * x[2]
            rb[2, 0] = 2.0 * x[0];          rb[2, 1] = 2.0 * x[1];          rb[2, 2] = 2.0 * x[2]
            rb[3, 0] = 2.0 * (x[0] - 1.0);  rb[3, 1] = 2.0 * x[1];          rb[3, 2] = 2.0 * x[2]


class HDIVBDM1(FEType):
    """src/fedefs/hdiv_bdm1.jl"""
    code = FE_HDIVBDM1
    coefficients = True
    hdiv = True

    def __init__(self, edim=2):
        self.edim = edim
        self.ncomponents = edim
        self.name = f"HDIVBDM1{{{edim}}}"

    def ndofs(self, edim):            # hdiv_bdm1.jl:20-23
        return edim + 1 if edim < self.edim else edim * _nn(edim)

    def ndofs_all(self, edim):
        if edim < self.edim:
            return edim + 1
        return 2 * _nn(edim) if edim == 2 else 4 * _nn(edim)

    def ncomponents_on(self, edim):
        return 1 if edim < self.edim else self.ncomponents

    def polynomialorder(self, edim):
        return 1

    def dofmap_pattern(self, edim):
        return "f2" if edim == 2 else "f3"

    def bface_dofmap_pattern(self, fdim):   # hdiv_bdm1.jl:33, 36
        return "i2" if fdim == 1 else "i3"

    def basis(self, rb, x, edim):
        if edim < self.edim:          # normal-flux face bases, hdiv_bdm1.jl:74-79, 109-115
            rb[0, 0] = 1.0
            if edim == 1:
                rb[1, 0] = 12.0 * (x[0] - 0.5)
            else:
                rb[1, 0] = 12.0 * (2.0 * x[0] + x[1] - 1.0)
                rb[2, 0] = 12.0 * (2.0 * x[1] + x[0] - 1.0)
        elif edim == 2:
            rb[0, 0] = x[0];        rb[0, 1] = x[1] - 1.0
            rb[2, 0] = x[0];        rb[2, 1] = x[1]
            rb[4, 0] = x[0] - 1.0;  rb[4, 1] = x[1]
            rb[1, 0] = 6.0 * x[0];                        rb[1, 1] = 6.0 - 12.0 * x[0] - 6.0 * x[1]
            rb[3, 0] = -6.0 * x[0];                       rb[3, 1] = 6.0 * x[1]
            rb[5, 0] = 6.0 * (x[0] - 1.0) + 12.0 * x[1];  rb[5, 1] = -6.0 * x[1]
        else:
            z = Dual(0.0) if isinstance(x[0], Dual) else 0.0
            rb[0, 0] = 2.0 * x[0];           rb[0, 1] = 2.0 * x[1];           rb[0, 2] = 2.0 * (x[2] - 1.0)
            rb[4, 0] = 2.0 * x[0];           rb[4, 1] = 2.0 * (x[1] - 1.0);   rb[4, 2] = 2.0 * x[2]
            rb[8, 0] = 2.0 * x[0];           rb[8, 1] = 2.0 * x[1];           rb[8, 2] = 2.0 * x[2]
            rb[12, 0] = 2.0 * (x[0] - 1.0);  rb[12, 1] = 2.0 * x[1];          rb[12, 2] = 2.0 * x[2]
            rb.last = 1.0 - x[0] - x[1] - x[2]
            l = rb.last
            rb[1, 0] = 24.0 * x[0];   rb[1, 1] = z;              rb[1, 2] = 24.0 * (l - x[0])
            rb[2, 0] = z;             rb[2, 1] = -24.0 * x[1];   rb[2, 2] = -24.0 * (l - x[1])
            rb[3, 0] = -24.0 * x[0];  rb[3, 1] = 24.0 * x[1];    rb[3, 2] = -24.0 * (x[1] - x[0])
            rb[5, 0] = z;             rb[5, 1] = 24.0 * (l - x[2]);     rb[5, 2] = 24.0 * x[2]
            rb[6, 0] = -24.0 * x[0];  rb[6, 1] = -24.0 * (l - x[0]);    rb[6, 2] = z
            rb[7, 0] = 24.0 * x[0];   rb[7, 1] = -24.0 * (x[0] - x[2]); rb[7, 2] = -24.0 * x[2]
            rb[9, 0] = -24.0 * x[0];  rb[9, 1] = z;              rb[9, 2] = 24.0 * x[2]
            rb[10, 0] = 24.0 * x[0];  rb[10, 1] = -24.0 * x[1];  rb[10, 2] = z
            rb[11, 0] = z;            rb[11, 1] = 24.0 * x[1];   rb[11, 2] = -24.0 * x[2]
            rb[13, 0] = 24.0 * (l - x[1]);     rb[13, 1] = 24.0 * x[1];   rb[13, 2] = z
            rb[14, 0] = -24.0 * (l - x[2]);    rb[14, 1] = z;             rb[14, 2] = -24.0 * x[2]
            rb[15, 0] = -24.0 * (x[2] - x[1]); rb[15, 1] = -24.0 * x[1];  rb[15, 2] = 24.0 * x[2]


class L2P0(FEType):
    """src/fedefs/l2_p0.jl -- piecewise constants, always broken (finiteelements.jl:78-80)"""
    code = FE_L2P0
    broken = True

    def __init__(self, ncomponents=1):
        self.ncomponents = ncomponents
        self.name = f"L2P0{{{ncomponents}}}"

    def ndofs(self, edim):
        return self.ncomponents

    def polynomialorder(self, edim):
        return 0

    def dofmap_pattern(self, edim):
        return "I1"

    def basis(self, rb, x, edim):
        for k in range(self.ncomponents):
            rb[k, k] = Dual(1.0) if isinstance(x[0], Dual) else 1.0


def reference_tables(fetype: FEType, edim: int, xref: np.ndarray, derivatives: bool):
    """refbasisvals / refbasisderivvals at the quadrature points (feevaluator.jl:64-68, 235-293).

    Returns (values[nq, nd_all, ncomp], derivs[nq, edim, nd_all*ncomp] or None); these
    are laid out exactly as include/grmp.h expects them."""
    nq = xref.shape[0]
    nda = fetype.ndofs_all(edim)
    nc = fetype.ncomponents_on(edim) if hasattr(fetype, "ncomponents_on") else fetype.ncomponents
    vals = np.zeros((nq, nda, nc))
    der = np.zeros((nq, edim, nda * nc)) if derivatives else None
    for i in range(nq):
        rb = _RefB(nda, nc, 0.0)
        fetype.basis(rb, [float(v) for v in xref[i]], edim)
        for d in range(nda):
            for c in range(nc):
                vals[i, d, c] = rb[d, c]
        if derivatives:
            xs = []
            for j in range(edim):
                e = np.zeros(3)
                e[j] = 1.0
                xs.append(Dual(xref[i, j], e))
            rbd = _RefB(nda, nc, Dual(0.0))
            fetype.basis(rbd, xs, edim)
            for c in range(nc):
                for d in range(nda):
                    v = rbd[d, c]
                    if isinstance(v, Dual):
                        der[i, :, d + c * nda] = v.d[:edim]
    return vals, der
