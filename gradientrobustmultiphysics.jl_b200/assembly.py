"""AssemblyPattern / assemble! (host mirror of src/assemblypatterns.jl,
src/assemblypatterns/bilinearform.jl, src/assemblypatterns/linearform.jl).

Same names, argument meaning and error behaviour as the reference; the cell loop itself
runs in libgrmp_cuda (include/grmp.h) -- there is no CPU fallback:

  DiscreteBilinearForm / DiscreteSymmetricBilinearForm / DiscreteLumpedBilinearForm
                                                (bilinearform.jl:38-75)
  DiscreteLinearForm                            (linearform.jl:29-33)
  prepare_assembly(AP)                          (assemblypatterns.jl:467-671: quadrature order
                                                 bonus + sum(polyorder + shift), evaluator tables)
  assemble(A, AP; factor, transposed_assembly, transpose_copy, skip_preps)   (bilinearform.jl:92-400)
  assemble(b, AP; factor, offset, skip_preps)                                (linearform.jl:47-251)
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib
from .fedefs import FEType, HDIVBDM1, HDIVRT0, H1BR, reference_tables
from .fespace import FEMatrix, FEMatrixBlock, FESpace, FEVectorBlock
from .quadrature import QuadratureRule

# ---- function operators (src/functionoperators.jl:13-153) ---------------------------------------


class AbstractFunctionOperator:
    code = 0
    needed_derivative = 0          # NeededDerivative4Operator (206-225)
    name = "??"

    def __repr__(self):
        return self.name

    def __eq__(self, o):
        return type(self) is type(o) and self.__dict__ == o.__dict__

    def __hash__(self):
        return hash(type(self).__name__)


class _Identity(AbstractFunctionOperator):
    code, name = 1, "id"


class _Gradient(AbstractFunctionOperator):
    code, needed_derivative, name = 2, 1, "∇"


class _SymmetricGradient(AbstractFunctionOperator):
    """SymmetricGradient{offdiagval}; only offdiagval = 1 is on the ported path (pdeoperators.jl:260)"""
    code, needed_derivative, name = 3, 1, "ϵ"

    def __init__(self, offdiagval=1):
        if offdiagval != 1:
            raise NotImplementedError("SymmetricGradient{offdiagval != 1} is not on the ported path")
        self.offdiagval = offdiagval


class _Divergence(AbstractFunctionOperator):
    code, needed_derivative, name = 4, 1, "div"


class ReconstructionIdentity(AbstractFunctionOperator):
    """ReconstructionIdentity{FETypeReconst} (feevaluator.jl:142-217)"""
    name = "R"

    def __init__(self, FETypeReconst: FEType):
        if not isinstance(FETypeReconst, (HDIVRT0, HDIVBDM1)):
            raise NotImplementedError("ReconstructionIdentity is ported for HDIVRT0 / HDIVBDM1 targets")
        self.FETypeReconst = FETypeReconst
        self.code = 5 if isinstance(FETypeReconst, HDIVRT0) else 6

    def __hash__(self):
        return hash(("R", self.code))


class _NormalFlux(AbstractFunctionOperator):
    """NormalFlux (functionoperators.jl:49): v_h . n_F of an Hdiv function on (boundary) faces, feevaluator_hdiv.jl:42-50"""
    code, name = 7, "NormalFlux"


Identity = _Identity()
NormalFlux = _NormalFlux()
Gradient = _Gradient()
SymmetricGradient = _SymmetricGradient
Divergence = _Divergence()


def _op(o):
    return o() if isinstance(o, type) else o


# ---- actions (src/actions.jl) -------------------------------------------------------------------
class NoAction:
    code = 0
    params = None

    def __init__(self, name="no action", bonus_quadorder=0):
        self.name, self.bonus_quadorder = name, bonus_quadorder


class HookeAction:
    """the tensor_apply_2d / tensor_apply_3d kernels of HookStiffnessOperator2D/3D
    (pdeoperators.jl:265-270, 304-312), evaluated on the device"""

    def __init__(self, dim, mu, lam):
        self.code = 1 if dim == 2 else 2
        self.params = np.array([mu, lam], dtype=np.float64)
        self.argsizes = [3, 3] if dim == 2 else [6, 6]
        self.bonus_quadorder = 0
        self.name = "hooke tensor"


class Action:
    """Action(kernel, argsizes; ...) with a user closure (actions.jl:53-63) cannot cross the C ABI."""

    def __init__(self, *a, **k):
        raise NotImplementedError("arbitrary Action kernels are host closures; only NoAction and the Hooke tensors run on the device "
                                  "(SURVEY.md 7, hard part 4). Use DataFunction for linear-form data.")


class DataFunction:
    """DataFunction(f, argsizes; dependencies, bonus_quadorder) (userdata.jl:35-63); `f` may be a
    constant vector or a callable x -> values evaluated by the host at the quadrature points."""

    def __init__(self, f, argsizes=None, dependencies="", bonus_quadorder=0, name="user data"):
        self.name, self.bonus_quadorder = name, bonus_quadorder
        if callable(f):
            self.kernel, self.constant = f, None
            self.argsizes = argsizes
            self.dependencies = dependencies or "X"
        else:
            self.constant = np.atleast_1d(np.asarray(f, dtype=np.float64))
            self.kernel = None
            self.argsizes = [self.constant.size, 0]
            self.dependencies = ""


class ConvectionAction:
    """kernel of ConvectionOperator(a_from, a_operator, xdim, ncomponents; a_to = 1) (pdeoperators.jl:459-468):
    result[j] = sum_k input[k] * input[xdim + (j-1) xdim + k], input = [a(x), ansatz operator evaluation]; evaluated on the device
    (GRMP_ACT_CONVECTION) with the coefficient function a as the fixed argument of a trilinear form"""
    code = 3
    params = None

    def __init__(self, xdim, ncomponents, bonus_quadorder=0, name="convection"):
        self.xdim, self.ncomponents, self.bonus_quadorder, self.name = xdim, ncomponents, bonus_quadorder, name
        self.argsizes = [ncomponents, xdim + ncomponents * xdim]


class NewtonConvectionAction:
    """OperatorWithUserJacobian(convection_function_fe_1, convection_jacobian, argsizes) of ConvectionOperator(...; newton = true)
    (pdeoperators.jl:459-493), evaluated on the device (GRMP_ACT_NEWTON_CONVECTION)"""
    code = 4
    params = None

    def __init__(self, xdim, ncomponents, bonus_quadorder=0, name="convection [Newton]"):
        self.xdim, self.ncomponents, self.bonus_quadorder, self.name = xdim, ncomponents, bonus_quadorder, name
        self.argsizes = [ncomponents, xdim + ncomponents * xdim, xdim + ncomponents * xdim]


class _FDotAction:
    """fdot_action(data) (actions.jl:119-128)"""
    code = 0

    def __init__(self, data: DataFunction):
        self.data = data
        self.bonus_quadorder = data.bonus_quadorder
        self.name = data.name


def fdot_action(data):
    return _FDotAction(data)


class _FDotNAction(_FDotAction):
    """fdotn_action(data, xgrid; bfaces = true) (actions.jl:175-192): f(x) . n_F of the boundary face, one component"""

    def __init__(self, data: DataFunction, xgrid):
        super().__init__(data)
        self.xgrid = xgrid
        self.name = data.name + "⋅n"


def fdotn_action(data, xgrid, bfaces=True):
    if not bfaces:
        raise NotImplementedError("fdotn_action on all faces belongs to ON_FACES assembly (not on the device)")
    return _FDotNAction(data, xgrid)


# numeric back end requested for newly prepared bilinear forms (include/grmp.h GRMP_PATH_*); tests pin the bit-exact path here
DEFAULT_PATH = _lib.PATH_AUTO


def _release(fn_name, handle):
    """finalizer of a device object: the library owns the memory behind the handle (grmp.h 'Ownership')"""
    try:
        getattr(_lib.lib(), fn_name)(handle)
    except Exception:
        pass


# ---- device mirrors of grid / space ---------------------------------------------------------------
def device_grid(xgrid, need_faces=False):
    L = _lib.lib()
    d = getattr(xgrid, "_dev", None)
    if d is None:
        h = C.c_void_p()
        vol = np.ascontiguousarray(xgrid.cellvolumes)
        if getattr(xgrid, "embedded", False):      # ON_BFACES: BFaceNodes / BFaceVolumes / BFaceRegions as the items
            _lib.check(L.grmp_grid_create_bfaces(_lib.context(), xgrid.xdim, xgrid.nnodes, _lib.ptr(xgrid.coords), xgrid.ncells,
                                                 _lib.ptr(xgrid.cellnodes), _lib.ptr(vol), _lib.ptr(xgrid.cellregions), C.byref(h)))
        else:
            _lib.check(L.grmp_grid_create(_lib.context(), xgrid.dim, xgrid.nnodes, _lib.ptr(xgrid.coords), xgrid.ncells,
                                          _lib.ptr(xgrid.cellnodes), _lib.ptr(vol), _lib.ptr(xgrid.cellregions), C.byref(h)))
        d = {"h": h, "faces": False}
        xgrid._dev = d
        weakref.finalize(xgrid, _release, "grmp_grid_destroy", h)
    if need_faces and not d["faces"]:
        ori = np.ascontiguousarray(xgrid.cellfaceorientations) if xgrid.dim == 3 else None
        _lib.check(L.grmp_grid_set_faces(d["h"], xgrid.nfaces, _lib.ptr(xgrid.cellfaces), _lib.ptr(xgrid.cellfacesigns),
                                         _lib.ptr(ori), _lib.ptr(xgrid.facenormals), _lib.ptr(xgrid.facevolumes)))
        d["faces"] = True
    return d["h"]


def device_space(FES):
    need_faces = FES.fetype.code in (3, 4, 5) and not getattr(FES.xgrid, "embedded", False)    # face bases carry no coefficients
    gh = device_grid(FES.xgrid, need_faces)
    d = getattr(FES, "_dev", None)
    if d is None:
        h = C.c_void_p()
        dofs = FES.celldofs
        _lib.check(_lib.lib().grmp_space_create(gh, FES.fetype.code, FES.fetype.ncomponents, FES.ndofs, dofs.shape[1], _lib.ptr(dofs),
                                                C.byref(h)))
        FES._dev = d = h
        # the space handle refers to the grid handle: keep the grid alive as long as the space, destroy the space first
        weakref.finalize(FES, _release, "grmp_space_destroy", h)
    return d


# ---- assembly patterns ---------------------------------------------------------------------------
APT_BilinearForm, APT_SymmetricBilinearForm, APT_LumpedBilinearForm, APT_LinearForm, APT_ItemIntegrator, APT_NonlinearForm = 0, 1, 2, 10, 20, 30


class AssemblyPattern:
    """AssemblyPattern{APT,T,AT} (assemblypatterns.jl:312-326), AT = ON_CELLS, or ON_BFACES for Identity forms of H1P1 / H1P2"""

    def __init__(self, APT, name, FES, operators, action, apply_action_to, regions, AT="ON_CELLS"):
        if AT not in ("ON_CELLS", "ON_BFACES"):
            raise NotImplementedError(f"assembly type {AT}: ON_CELLS and ON_BFACES are on the device")
        self.APT, self.name, self.FES, self.AT = APT, name, list(FES), AT
        self.operators = [_op(o) for o in operators]
        self.action, self.apply_action_to, self.regions = action, apply_action_to, list(regions)
        self.last_allocations = 0
        self.AM = None                      # prepared state (quadrature, tables, device handle)
        self.fixed = None                   # (FESpace, operator) of the coefficient argument of a trilinear form (FES[1] of nFE = 3)

    def item_space(self, i):
        """the space as the item loop sees it: ItemDofs = Dofmap4AssemblyType(FES, AT) (assemblypatterns.jl:414-420)"""
        F = self.FES[i]
        return F.on_bfaces() if self.AT == "ON_BFACES" else F

    def __repr__(self):
        return f"AssemblyPattern({self.name}, {self.FES}, {self.operators})"


def DiscreteBilinearForm(operators, FES, action=None, name="BLF", regions=(0,), apply_action_to=(1,), AT="ON_CELLS"):
    assert len(operators) == len(FES), "each FESpace needs an operator and vice versa"
    if len(FES) == 3:
        assert AT == "ON_CELLS"
        # trilinear form with one coefficient argument: FES = [FES_a, FES_ansatz, FES_test] (bilinearform.jl:60-64, 235-257)
        if not isinstance(action, ConvectionAction):
            raise NotImplementedError("trilinear forms: the convection kernel runs on the device; other actions are user closures")
        AP = AssemblyPattern(APT_BilinearForm, name, FES[1:], operators[1:], action, [1], regions)
        AP.fixed = (FES[0], _op(operators[0]))
        return AP
    assert len(FES) == 2, "bilinear forms take two FESpaces (+ one coefficient argument)"
    assert list(apply_action_to) == [1], "the ported path applies the action to argument 1 (all operators on the path do)"
    return AssemblyPattern(APT_BilinearForm, name, FES, operators, action or NoAction(), [1], regions, AT)


def DiscreteSymmetricBilinearForm(operators, FES, action=None, name="symBLF", regions=(0,), apply_action_to=(1,), AT="ON_CELLS"):
    assert len(operators) == len(FES) == 2, "each FESpace needs an operator and vice versa"
    return AssemblyPattern(APT_SymmetricBilinearForm, name, FES, operators, action or NoAction(), [1], regions, AT)


def DiscreteLumpedBilinearForm(operators, FES, action=None, name="lumpedBLF", regions=(0,), apply_action_to=(1,), AT="ON_CELLS"):
    assert len(operators) == len(FES) == 2, "each FESpace needs an operator and vice versa"
    return AssemblyPattern(APT_LumpedBilinearForm, name, FES, operators, action or NoAction(), [1], regions, AT)


def DiscreteNonlinearForm(operators, FES, action, name="NLF", regions=(0,)):
    """AssemblyPattern{APT_NonlinearForm} (nonlinearform.jl:1-40) for the Newton convection form: operators = [a_operator,
    ansatz_operator, test_operator], all three FESpaces the space of the unknown (nonlinearform.jl:110-112)"""
    assert len(operators) == len(FES) == 3 and FES[0] is FES[1] is FES[2], "the nonlinearity depends on one unknown"
    if not isinstance(action, NewtonConvectionAction):
        raise NotImplementedError("NonlinearForm: the Newton convection kernel runs on the device; other kernels are user closures")
    AP = AssemblyPattern(APT_NonlinearForm, name, FES[1:], operators[1:], action, [1], regions)
    AP.fixed = (FES[0], _op(operators[0]))
    return AP


def full_assemble(A, b, AP: AssemblyPattern, FEB, factor=1, transposed_assembly=False, skip_preps=False):
    """full_assemble!(A, b, AP, FEB; factor, transposed_assembly, skip_preps) (nonlinearform.jl:44-245): Jacobian into the matrix
    block A, DN(u) u - N(u) into the vector block b (may be None)"""
    assert AP.APT == APT_NonlinearForm and isinstance(A, FEMatrixBlock)
    blk = FEB[0] if isinstance(FEB, (list, tuple)) else FEB
    assert blk.FES is AP.fixed[0]
    tr = bool(transposed_assembly)
    if AP.AM is None or AP.AM.kind != "blf" or AP.AM.transposed != tr:
        prepare_assembly(AP, tr)
    P = AP.AM
    coeffs = np.ascontiguousarray(blk.entries[blk.offset:blk.offset + blk.FES.ndofs])
    _lib.check(_lib.lib().grmp_blf_set_newton_argument(P.h, AP.fixed[1].code, C.byref(P.fixed_tab), _lib.ptr(coeffs),
                                                       int(bool(skip_preps and P.have_pattern))))
    cp, rv, nz = assemble_csc(AP, factor, skip_preps, transposed_assembly)
    A.parent.add_csc(*_embed(A, cp, rv, nz))
    if b is not None:
        assert isinstance(b, FEVectorBlock) and b.FES is AP.FES[1]
        _lib.check(_lib.lib().grmp_blf_newton_rhs(P.h, _lib.ptr(b.entries), int(b.offset)))
    AP.last_allocations = 0
    return None


def DiscreteLinearForm(operators, FES, action=None, name="LF", regions=(0,), AT="ON_CELLS"):
    assert len(operators) == len(FES), "each FESpace needs an operator and vice versa"
    if len(FES) == 2:
        assert AT == "ON_CELLS"
        # one coefficient argument, NoAction: FES = [FES_a, FES_test] (linearform.jl:29-33, 130-178)
        if action is not None and not isinstance(action, NoAction):
            raise NotImplementedError("LinearForms with a coefficient argument: NoAction runs on the device, user actions stay with the reference")
        AP = AssemblyPattern(APT_LinearForm, name, FES[1:], operators[1:], NoAction(), [1], regions)
        AP.fixed = (FES[0], _op(operators[0]))
        return AP
    if len(FES) != 1:
        raise NotImplementedError("LinearForms with several coefficient arguments are a 'next' row (SURVEY.md 8f N4)")
    return AssemblyPattern(APT_LinearForm, name, FES, operators, action or NoAction(), [1], regions, AT)


def _geometry(xgrid):
    return {1: "Edge1D", 2: "Triangle2D", 3: "Tetrahedron3D"}[xgrid.dim]


def _tables(FES, op, qf):
    """reference tables of FEEvaluator(FES, op, qf) as a grmp_evaltab (+ arrays kept alive)"""
    edim = FES.xgrid.dim
    fe = op.FETypeReconst if isinstance(op, ReconstructionIdentity) else FES.fetype
    vals, der = reference_tables(fe, edim, qf.xref, op.needed_derivative > 0)
    vals = np.ascontiguousarray(vals)
    der = None if der is None else np.ascontiguousarray(der)
    tab = _lib.EvalTab(fe.ndofs_all(edim), vals.shape[2], _lib.ptr(vals), _lib.ptr(der))
    return tab, (vals, der)


class _Prepared:
    """prepared state of an AssemblyPattern; owns the device-side pattern object"""

    def __del__(self):
        h, kind = getattr(self, "h", None), getattr(self, "kind", None)
        if h is not None and kind is not None:
            _release({"blf": "grmp_blf_destroy", "lf": "grmp_lf_destroy", "ii": "grmp_ii_destroy"}[kind], h)
            self.h = None


def quadrature_order(AP: AssemblyPattern):
    """assemblypatterns.jl:559-565"""
    edim = AP.item_space(0).xgrid.dim
    q = AP.action.bonus_quadorder
    for F, o in zip(AP.FES, AP.operators):
        q += F.fetype.polynomialorder(edim) - o.needed_derivative
    if getattr(AP, "fixed", None) is not None:          # every FESpace of the pattern counts, coefficient arguments included
        q += AP.fixed[0].fetype.polynomialorder(edim) - AP.fixed[1].needed_derivative
    return max(q, 0)


def prepare_assembly(AP: AssemblyPattern, transposed_assembly=False):
    """prepare_assembly!(AP): quadrature rule, evaluator tables, device-side pattern object"""
    L = _lib.lib()
    P = _Prepared()
    xgrid = AP.item_space(0).xgrid
    P.quadorder = quadrature_order(AP)
    P.qf = QuadratureRule(_geometry(xgrid), P.quadorder)
    P.keep = []
    regions = np.ascontiguousarray(AP.regions, dtype=np.int32)
    w = np.ascontiguousarray(P.qf.w)
    h = C.c_void_p()
    if AP.APT == APT_ItemIntegrator:
        sp = device_space(AP.item_space(0))
        tab, keep = _tables(AP.item_space(0), AP.operators[0], P.qf)
        P.keep.append(keep)
        _lib.check(L.grmp_ii_create(sp, AP.operators[0].code, AP.action.code, _lib.ptr(regions), regions.size, len(P.qf), _lib.ptr(w),
                                    C.byref(tab), C.byref(h)))
        P.kind = "ii"
    elif AP.APT == APT_LinearForm:
        sp = device_space(AP.item_space(0))
        tab, keep = _tables(AP.item_space(0), AP.operators[0], P.qf)
        P.keep.append(keep)
        _lib.check(L.grmp_lf_create(sp, AP.operators[0].code, _lib.ptr(regions), regions.size, len(P.qf), _lib.ptr(w), C.byref(tab),
                                    C.byref(h)))
        P.kind = "lf"
        if DEFAULT_PATH != _lib.PATH_AUTO:
            _lib.check(L.grmp_lf_set_path(h, _lib.PATH_GENERIC if DEFAULT_PATH == _lib.PATH_GENERIC else _lib.PATH_COLUMNS))
    else:
        s1, s2 = device_space(AP.item_space(0)), device_space(AP.item_space(1))
        t1, k1 = _tables(AP.item_space(0), AP.operators[0], P.qf)
        t2, k2 = _tables(AP.item_space(1), AP.operators[1], P.qf)
        P.keep += [k1, k2]
        act = AP.action
        if not isinstance(act, (NoAction, HookeAction, ConvectionAction, NewtonConvectionAction)):
            raise NotImplementedError("bilinear forms support NoAction, the Hooke tensor actions and the convection kernels on the device")
        if isinstance(act, (ConvectionAction, NewtonConvectionAction)):
            P.fixed_tab, kf = _tables(AP.fixed[0], AP.fixed[1], P.qf)
            P.keep.append(kf)
        _lib.check(L.grmp_blf_create(s1, s2, AP.operators[0].code, AP.operators[1].code, act.code, _lib.ptr(act.params), 0 if AP.APT == APT_NonlinearForm else AP.APT,
                                     int(bool(transposed_assembly)), _lib.ptr(regions), regions.size, len(P.qf), _lib.ptr(w), C.byref(t1), C.byref(t2), C.byref(h)))
        P.kind = "blf"
        if DEFAULT_PATH != _lib.PATH_AUTO:
            _lib.check(L.grmp_blf_set_path(h, DEFAULT_PATH))
        P.have_pattern = False
        P.transposed = bool(transposed_assembly) and AP.APT != APT_SymmetricBilinearForm
    P.h = h
    AP.AM = P
    return P


def blf_set_path(AP: AssemblyPattern, path: int):
    if AP.AM is None:
        prepare_assembly(AP)
    _lib.check(_lib.lib().grmp_blf_set_path(AP.AM.h, path))


def blf_stats(AP):
    st = _lib.Stats()
    fn = _lib.lib().grmp_blf_stats if AP.AM.kind == "blf" else _lib.lib().grmp_lf_stats
    _lib.check(fn(AP.AM.h, C.byref(st)))
    return st


def assemble_csc(AP: AssemblyPattern, factor=1.0, skip_preps=False, transposed_assembly=False, fetch=True):
    """numeric core of assemble!(A, AP): returns the operator's own CSC (colptr, rowval, nzval),
    1-based Int64 -- SparseMatrixCSC{Float64,Int64} of a fresh ExtendableSparseMatrix after flush!."""
    L = _lib.lib()
    tr = bool(transposed_assembly) and AP.APT != APT_SymmetricBilinearForm
    if AP.AM is None or AP.AM.kind != "blf" or AP.AM.transposed != tr:
        prepare_assembly(AP, tr)        # transposed_assembly is a creation-time property of the device pattern
    P = AP.AM
    if not P.have_pattern or not skip_preps:
        nnz = C.c_int64(0)
        _lib.check(L.grmp_blf_symbolic(P.h, float(factor), C.byref(nnz)))
        P.nnz = nnz.value
        ncols = AP.FES[0].ndofs if P.transposed else AP.FES[1].ndofs
        P.colptr = np.zeros(ncols + 1, np.int64)
        P.rowval = np.zeros(P.nnz, np.int64)
        _lib.check(L.grmp_blf_get_pattern(P.h, _lib.ptr(P.colptr), _lib.ptr(P.rowval)))
        P.have_pattern = True
    nzval = np.zeros(P.nnz) if fetch else None
    _lib.check(L.grmp_blf_numeric(P.h, float(factor), _lib.ptr(nzval)))
    return P.colptr, P.rowval, nzval


# ---- the matrix stays on the device: products, residuals, penalties --------------------------------
def addblock_matmul(a: np.ndarray, AP: AssemblyPattern, b: np.ndarray, factor=1, transposed=False):
    """addblock_matmul!(a, B, b; factor, transposed) (fematrix.jl:402-473) with B = the device-resident matrix of the last
    assemble_csc(AP, ...): a += B*b*factor (or B'*b*factor), bit-identical to the reference loop"""
    assert a.dtype == np.float64 and b.dtype == np.float64 and a.flags.c_contiguous and b.flags.c_contiguous
    _lib.check(_lib.lib().grmp_blf_matmul(AP.AM.h, _lib.ptr(b), _lib.ptr(a), float(factor), int(bool(transposed))))
    return a


def residual(AP: AssemblyPattern, x: np.ndarray, b: np.ndarray | None = None, fixed_dofs=None, want_vector=True):
    """residual check of solve_direct! (solvers.jl:661-668): r = A*x - b, r[fixed_dofs] = 0; returns (r, sum r_i^2)"""
    nrows = AP.FES[1].ndofs if AP.AM.transposed else AP.FES[0].ndofs
    r = np.zeros(nrows) if want_vector else None
    fd = None if fixed_dofs is None else np.ascontiguousarray(fixed_dofs, dtype=np.int64)
    nrm = C.c_double(0)
    _lib.check(_lib.lib().grmp_blf_residual(AP.AM.h, _lib.ptr(np.ascontiguousarray(x)), _lib.ptr(b), _lib.ptr(fd), 0 if fd is None else fd.size,
                                            _lib.ptr(r), C.byref(nrm)))
    return r, nrm.value


def apply_penalties(AP: AssemblyPattern, fixed_dofs, penalty):
    """apply_penalties!(A, fixed_dofs, penalty) (fematrix.jl:349-355) on the device-resident values"""
    fd = np.ascontiguousarray(fixed_dofs, dtype=np.int64)
    miss = C.c_int64(0)
    _lib.check(_lib.lib().grmp_blf_apply_penalties(AP.AM.h, _lib.ptr(fd), fd.size, float(penalty), C.byref(miss)))


def device_csc(AP: AssemblyPattern):
    """device pointers of the assembled SparseMatrixCSC (hand-off to a GPU solver, solvers.jl:655)"""
    d = _lib.DeviceCSC()
    _lib.check(_lib.lib().grmp_blf_device_csc(AP.AM.h, C.byref(d)))
    return d


def fetch_values(AP: AssemblyPattern):
    nz = np.zeros(AP.AM.nnz)
    _lib.check(_lib.lib().grmp_blf_get_values(AP.AM.h, _lib.ptr(nz)))
    return nz


def _embed(block: FEMatrixBlock, colptr, rowval, nzval):
    """place a block CSC at (offsetX, offsetY) of the parent matrix"""
    par = block.parent
    cp = np.full(par.n + 1, 1, dtype=np.int64)
    counts = np.zeros(par.n, np.int64)
    counts[block.offsetY:block.last_indexY] = np.diff(colptr)
    cp[1:] = 1 + np.cumsum(counts)
    return cp, rowval + block.offsetX, nzval


def assemble(target, AP: AssemblyPattern, FEB=(), factor=1, factor_transpose=None, transposed_assembly=False, transpose_copy=None,
             skip_preps=False, fixed_arguments=None, offset=0, fdata=None):
    """assemble!(A::FEMatrixBlock, AP; ...) / assemble!(b::FEVectorBlock | Vector, AP; ...)"""
    if AP.APT == APT_LinearForm and AP.fixed is not None:
        assert len(FEB) == 1 and FEB[0].FES is AP.fixed[0], "LinearForm with a coefficient argument: FEB = [block of the coefficient function]"
        return _assemble_lf(target, AP, factor=factor, skip_preps=skip_preps, offset=offset, feb=FEB[0])
    if len(FEB) != 0 and not (AP.fixed is not None and len(FEB) == 1 and AP.APT == APT_BilinearForm):
        raise NotImplementedError("FEB coefficient arguments: one fixed argument of a trilinear convection form is on the device "
                                  "(SURVEY.md 8f N4); anything else stays with the reference")
    if AP.fixed is not None:
        assert len(FEB) == 1 and FEB[0].FES is AP.fixed[0], "trilinear form: FEB = [block of the coefficient function]"
        tr = bool(transposed_assembly)
        if AP.AM is None or AP.AM.kind != "blf" or AP.AM.transposed != tr:
            prepare_assembly(AP, tr)
        blk = FEB[0]
        coeffs = np.ascontiguousarray(blk.entries[blk.offset:blk.offset + blk.FES.ndofs])
        _lib.check(_lib.lib().grmp_blf_set_fixed_argument(AP.AM.h, device_space(AP.fixed[0]), AP.fixed[1].code, C.byref(AP.AM.fixed_tab),
                                                          _lib.ptr(coeffs), int(bool(skip_preps and AP.AM.have_pattern))))
    if AP.APT == APT_LinearForm:
        return _assemble_lf(target, AP, factor=factor, skip_preps=skip_preps, offset=offset)
    assert isinstance(target, FEMatrixBlock), "assemble into an FEMatrixBlock (A[j,k])"
    tr = bool(transposed_assembly) and AP.APT != APT_SymmetricBilinearForm
    fx, fy = (AP.FES[1], AP.FES[0]) if tr else (AP.FES[0], AP.FES[1])
    assert target.FESX is fx and target.FESY is fy, "pattern spaces must match the block"
    if tr and transpose_copy is not None:
        raise NotImplementedError("transposed_assembly together with transpose_copy")
    cp, rv, nz = assemble_csc(AP, factor, skip_preps, transposed_assembly)
    target.parent.add_csc(*_embed(target, cp, rv, nz))
    if transpose_copy is not None:
        ft = factor if factor_transpose is None else factor_transpose
        L = _lib.lib()
        nrows = AP.FES[0].ndofs
        cpt = np.zeros(nrows + 1, np.int64)
        rvt = np.zeros(rv.size, np.int64)
        nzt = np.zeros(rv.size)
        _lib.check(L.grmp_blf_transpose_copy(AP.AM.h, float(factor), float(ft), _lib.ptr(cpt), _lib.ptr(rvt), _lib.ptr(nzt)))
        assert isinstance(transpose_copy, FEMatrixBlock)
        transpose_copy.parent.add_csc(*_embed(transpose_copy, cpt, rvt, nzt))
    AP.last_allocations = 0
    return None


def _qp_table(AP, P):
    """host evaluation of the DataFunction at x = b + A*xref (eval_trafo!, linearform.jl:197-201)"""
    data = AP.action.data
    g = AP.item_space(0).xgrid
    if isinstance(AP.action, _FDotNAction):     # f(x) . n_F per boundary face: item dependent, always a table
        pg = g.parent
        nrm = pg.facenormals[pg.bfacefaces.astype(np.int64) - 1]
        if data.constant is not None:
            vals = np.broadcast_to(data.constant, (g.ncells, len(P.qf), data.constant.size))
        else:
            vals = _qp_table_values(data, g, P)
        return 2, np.ascontiguousarray((vals * nrm[:, None, :]).sum(axis=2)[:, :, None])
    if data.constant is not None:
        return 1, data.constant
    return 2, _qp_table_values(data, g, P)


def _qp_table_values(data, g, P):
    x = g.coords
    cn = g.cellnodes.astype(np.int64) - 1
    b = x[cn[:, 0]]
    xq = np.repeat(b[:, None, :], len(P.qf), axis=1).copy()
    for j in range(g.dim):
        Aj = x[cn[:, j + 1]] - b                     # column j of A
        xq += Aj[:, None, :] * P.qf.xref[None, :, j, None]
    flat = xq.reshape(-1, x.shape[1])
    try:
        vals = np.asarray(data.kernel(flat.T), dtype=np.float64)       # vectorised: f(x[dim, npts]) -> [ncomp, npts]
        vals = vals.reshape(-1, flat.shape[0]).T
    except Exception:
        vals = np.array([np.atleast_1d(data.kernel(p)) for p in flat], dtype=np.float64)
    return np.ascontiguousarray(vals.reshape(g.ncells, len(P.qf), -1))


def _assemble_lf(b, AP, factor=1, skip_preps=False, offset=0, feb=None):
    L = _lib.lib()
    if AP.AM is None or not skip_preps:
        if AP.AM is None or AP.AM.kind != "lf":
            prepare_assembly(AP)
    P = AP.AM
    if isinstance(b, FEVectorBlock):
        assert b.FES is AP.FES[0]
        entries, offset = b.entries, b.offset
    else:
        entries = b
    assert entries.dtype == np.float64 and entries.flags.c_contiguous
    if feb is not None:
        if not hasattr(P, "fixed_tab"):
            P.fixed_tab, kf = _tables(AP.fixed[0], AP.fixed[1], P.qf)
            P.keep.append(kf)
        coeffs = np.ascontiguousarray(feb.entries[feb.offset:feb.offset + feb.FES.ndofs])
        _lib.check(L.grmp_lf_assemble_feb(P.h, float(factor), device_space(AP.fixed[0]), AP.fixed[1].code, C.byref(P.fixed_tab), _lib.ptr(coeffs),
                                          _lib.ptr(entries), int(offset)))
        AP.last_allocations = 0
        return None
    if isinstance(AP.action, _FDotAction):
        fsrc, fd = _qp_table(AP, P)
        rd = AP.operators[0]
        if fd.shape[-1] != _resultdim(AP):
            raise ValueError(f"data has {fd.shape[-1]} components, operator result has {_resultdim(AP)}")
    elif isinstance(AP.action, NoAction):
        fsrc, fd = 0, None
    else:
        raise NotImplementedError("linear forms support NoAction and fdot_action(DataFunction)")
    _lib.check(L.grmp_lf_assemble(P.h, float(factor), fsrc, _lib.ptr(fd), _lib.ptr(entries), int(offset)))
    AP.last_allocations = 0
    return None


# ---- ItemIntegrator (itemintegrator.jl) ----------------------------------------------------------------------------------------
class _IIAction:
    """the closed set of ItemIntegrator kernels evaluated on the device (grmp.h GRMP_II_*)"""

    def __init__(self, code, bonus_quadorder=0, data=None, factor=1.0, name="ItemIntegrator"):
        self.code, self.bonus_quadorder, self.data, self.factor, self.name = code, bonus_quadorder, data, float(factor), name


def ItemIntegrator(operators, action=None, regions=(0,), name="ItemIntegrator", AT="ON_CELLS"):
    """ItemIntegrator(operators, action; AT, regions) (itemintegrator.jl:18-21); one argument, NoAction or one of the
    integrators below; AT = ON_CELLS, or ON_BFACES for boundary integrals (Identity of H1P1 / H1P2, NormalFlux of HDIVRT0 / HDIVBDM1)"""
    if len(operators) != 1:
        raise NotImplementedError("ItemIntegrators with several arguments are a 'next' row (SURVEY.md 8f N4)")
    act = action if isinstance(action, _IIAction) else _IIAction(0)
    if action is not None and not isinstance(action, (_IIAction, NoAction)):
        raise NotImplementedError("user Actions cannot cross the C ABI: NoAction, L2NormIntegrator and L2ErrorIntegrator run on the device")
    return AssemblyPattern(APT_ItemIntegrator, name, [], operators, act, [1], regions, AT)


def L2NormIntegrator(ncomponents, operator, quadorder=2, regions=(0,), name="L2 norm", AT="ON_CELLS"):
    """L2NormIntegrator(ncomponents, operator; AT, quadorder = 2) (itemintegrator.jl:91-110)"""
    return ItemIntegrator([operator], _IIAction(1, bonus_quadorder=quadorder, name=name), regions=regions, name=name, AT=AT)


def L2ErrorIntegrator(compare_data: DataFunction, operator=None, quadorder="auto", factor=1, regions=(0,), name="auto", AT="ON_CELLS"):
    """L2ErrorIntegrator(compare_data, operator; quadorder = "auto", factor) (itemintegrator.jl:33-78): || compare_data - factor u_h ||^2
    per item; "auto" = twice the bonus quadrature order of the data"""
    q = 2 * compare_data.bonus_quadorder if quadorder == "auto" else int(quadorder)
    nm = f"L2 error ({compare_data.name})" if name == "auto" else name
    return ItemIntegrator([Identity if operator is None else operator], _IIAction(2, bonus_quadorder=q, data=compare_data, factor=factor, name=nm),
                          regions=regions, name=nm, AT=AT)


def _ii_prepare(AP, FEB, skip_preps):
    if isinstance(FEB, (list, tuple)):
        assert len(FEB) == 1
        FEB = FEB[0]
    assert isinstance(FEB, FEVectorBlock), "evaluate an ItemIntegrator on an FEVectorBlock"
    if AP.AM is None or not skip_preps or AP.FES != [FEB.FES]:
        AP.FES = [FEB.FES]                       # prepare_assembly!(AP, FE) (itemintegrator.jl:186-192)
        prepare_assembly(AP)
    P = AP.AM
    data = None
    if AP.action.code == 2:
        fsrc, fd = _qp_table(AP, P)                 # compare_data at the quadrature points (L2error_function, itemintegrator.jl:52-69)
        if fsrc == 1:
            fd = np.broadcast_to(fd, (AP.item_space(0).xgrid.ncells, len(P.qf), fd.size))
        data = np.ascontiguousarray(fd, dtype=np.float64)
        if data.shape[-1] != _resultdim(AP):
            raise ValueError(f"compare data has {data.shape[-1]} components, operator result has {_resultdim(AP)}")
    coeffs = np.ascontiguousarray(FEB.entries[FEB.offset:FEB.offset + FEB.FES.ndofs])
    rd = C.c_int(0)
    _lib.check(_lib.lib().grmp_ii_resultdim(P.h, C.byref(rd)))
    return P, coeffs, data, rd.value


def evaluate_itemwise(b, AP: AssemblyPattern, FEB, skip_preps=False):
    """evaluate!(b, AP, FEB) (itemintegrator.jl:160-300): b[item, j] += ... (Julia: b[j, item]); returns b"""
    P, coeffs, data, rd = _ii_prepare(AP, FEB, skip_preps)
    assert b.dtype == np.float64 and b.flags.c_contiguous and b.shape == (AP.item_space(0).xgrid.ncells, rd)
    _lib.check(_lib.lib().grmp_ii_evaluate(P.h, _lib.ptr(coeffs), AP.action.factor, _lib.ptr(data), _lib.ptr(b), None))
    return b


def evaluate(AP: AssemblyPattern, FEB, skip_preps=False):
    """evaluate(AP, FEB) (itemintegrator.jl:316-360): accumulation over all items; a scalar if the result has one component"""
    P, coeffs, data, rd = _ii_prepare(AP, FEB, skip_preps)
    tot = np.zeros(rd)
    _lib.check(_lib.lib().grmp_ii_evaluate(P.h, _lib.ptr(coeffs), AP.action.factor, _lib.ptr(data), None, _lib.ptr(tot)))
    return float(tot[0]) if rd == 1 else tot


def _resultdim(AP):
    F = AP.item_space(len(AP.FES) - 1) if hasattr(AP, "item_space") else AP.FES[-1]
    o = AP.operators[-1]
    edim, nc = F.xgrid.dim, F.fetype.ncomponents
    return {1: nc, 5: nc, 6: nc, 7: 1, 2: edim * nc, 3: (3 if edim == 2 else 6), 4: max(1, nc // edim)}[o.code]
