"""PDE operator front end (host mirror of src/pdeoperators.jl) -- the drop-in boundary.

  LaplaceOperator(κ)                      pdeoperators.jl:154-169  -> SymmetricBLF [∇, ∇], factor κ
  ReactionOperator(α)                     pdeoperators.jl:176-203  -> SymmetricBLF [id, id], factor α
  HookStiffnessOperator2D/3D(μ, λ)        pdeoperators.jl:256-315  -> BLF [ϵ, ϵ] with the Hooke tensor action
  LagrangeMultiplier(operator)            pdeoperators.jl:213-226  -> BLF [operator, id], factor -1, transposed copy
  BilinearForm(operators, action)         pdeoperators.jl:355-395
  LinearForm(operator, data | action)     pdeoperators.jl:659-705
  assemble_operator(A[j,k], O) / assemble_operator(b[j], O)      pdeoperators.jl:978-1006
"""
from __future__ import annotations

import numpy as np

from .assembly import (APT_BilinearForm, APT_LinearForm, APT_NonlinearForm, APT_SymmetricBilinearForm, AssemblyPattern, ConvectionAction,
                       NewtonConvectionAction, DiscreteNonlinearForm, full_assemble, DataFunction, Divergence,
                       Gradient, HookeAction, Identity, NoAction, SymmetricGradient, _FDotAction, _op, assemble, fdot_action)
from .fespace import FEMatrixBlock, FEVectorBlock


class PDEOperator:
    """PDEOperator{T,APT,AT} (pdeoperators.jl:63-100)"""

    def __init__(self, APT, name, operators4arguments, action, apply_action_to, factor, regions, store=False):
        self.APT, self.name = APT, name
        self.operators4arguments = [_op(o) for o in operators4arguments]
        self.action, self.apply_action_to = action, apply_action_to
        self.factor, self.regions, self.store_operator = factor, list(regions), store
        self.transposed_assembly = False
        self.transposed_copy = False
        self.transpose_factor = None
        self.fixed_arguments = []
        self.fixed_arguments_ids = []

    def __repr__(self):
        return f"PDEOperator({self.name})"


def LaplaceOperator(κ=1.0, name="auto", regions=(0,), store=False, **kw):
    if not isinstance(κ, (int, float)):
        raise ValueError("No standard Laplace operator definition for this type of κ available, please define your own action "
                         "and PDEOperator with it.")
    if name == "auto":
        name = ("" if κ == 1 else f"{κ} ") + "(∇#A,∇#T)"
    return PDEOperator(APT_SymmetricBilinearForm, name, [Gradient, Gradient], NoAction(), [1], κ, regions, store)


def ReactionOperator(α=1.0, ncomponents=1, name="auto", regions=(0,), store=False, **kw):
    if not isinstance(α, (int, float)):
        raise NotImplementedError("ReactionOperator with a DataFunction coefficient needs a user action kernel (not on the ported path)")
    if name == "auto":
        name = ("" if α == 1.0 else f"{α} ") + "(#A,#T)"
    return PDEOperator(APT_SymmetricBilinearForm, name, [Identity, Identity], NoAction(), [1], α, regions, store)


def LagrangeMultiplier(operator, name="auto", action=None, regions=(0,), store=False, factor=-1):
    if name == "auto":
        name = f"(#A, {operator}(#T))"
        name = ("-" + name) if factor == -1 else (name if factor == 1 else f"{factor} {name}")
    O = PDEOperator(APT_BilinearForm, name, [operator, Identity], action or NoAction(), [1], factor, regions, store)
    O.transposed_copy = True
    return O


def HookStiffnessOperator2D(μ, λ, name="(ℂ(μ,λ) ϵ(#A),ϵ(#T))", regions=(0,), ϵ=None, store=False):
    ϵ = ϵ or SymmetricGradient(1)
    return PDEOperator(APT_BilinearForm, name, [ϵ, ϵ], HookeAction(2, μ, λ), [1], 1, regions, store)


def HookStiffnessOperator3D(μ, λ, name="(ℂ(μ,λ) ϵ(#A),ϵ(#T))", regions=(0,), ϵ=None, store=False):
    ϵ = ϵ or SymmetricGradient(1)
    return PDEOperator(APT_BilinearForm, name, [ϵ, ϵ], HookeAction(3, μ, λ), [1], 1, regions, store)


def BilinearForm(operators_linear, action=None, name="auto", regions=(0,), factor=1, transposed_assembly=False,
                 also_transposed_block=False, transpose_factor=None, store=False, APT=APT_BilinearForm):
    """BilinearForm(operators_linear, action; ...) without `operators_current` (pdeoperators.jl:355-395)"""
    if name == "auto":
        name = f"({operators_linear[0]}(#A), {operators_linear[1]}(#T))"
    O = PDEOperator(APT, name, operators_linear, action or NoAction(), [1], factor, regions, store)
    O.transposed_assembly = transposed_assembly
    O.transposed_copy = also_transposed_block
    O.transpose_factor = transpose_factor
    return O


def LinearForm(operator, data=None, name="auto", regions=(0,), factor=1, store=False):
    """LinearForm(operator, f::DataFunction | action) (pdeoperators.jl:659-705)"""
    if isinstance(data, DataFunction):
        action = fdot_action(data)
        if name == "auto":
            name = f"({data.name}, {operator}(#T))"
    elif data is None or isinstance(data, (NoAction, _FDotAction)):
        action = data or NoAction()
        if name == "auto":
            name = f"(A({operator}(#T)), 1)"
    else:
        action = fdot_action(DataFunction(data))
        if name == "auto":
            name = f"(f, {operator}(#T))"
    O = PDEOperator(APT_LinearForm, name, [operator], action, [1], 1, regions, store)
    O.factor = factor
    return O


def ConvectionOperator(a_from: int, a_operator, xdim: int, ncomponents: int, name="auto", a_to=1, factor=1, ansatz_operator=Gradient,
                       test_operator=Identity, regions=(0,), newton=False, store=False, transposed_assembly=True, bonus_quadorder=0):
    """ConvectionOperator(a_from, a_operator, xdim, ncomponents; a_to = 1, ...) (pdeoperators.jl:435-510): the Picard-linearised
    convection term ((a . grad) u, v) as a trilinear form whose first argument is the coefficient function CurrentSolution[a_from]"""
    if newton:      # NonlinearForm(test_operator, [a_operator, ansatz_operator], [a_from, a_from], kernel, argsizes; jacobian) (pdeoperators.jl:481-493)
        if name == "auto":
            name = f"(({_op(a_operator)}(#1) . {_op(ansatz_operator)}) #1, {_op(test_operator)}(#T)) [Newton]"
        O = PDEOperator(APT_NonlinearForm, name, [a_operator, ansatz_operator, test_operator], NewtonConvectionAction(xdim, ncomponents, bonus_quadorder),
                        [1, 2], factor, regions, store)
        O.fixed_arguments = [1, 2]
        O.fixed_arguments_ids = [a_from, a_from]
        O.transposed_assembly = True
        return O
    if a_to != 1:
        raise NotImplementedError("ConvectionOperator: the coefficient in position 1 (a_to = 1) is on the device")
    if name == "auto":
        name = f"(({_op(a_operator)}(#1) . {_op(ansatz_operator)}) #A, {_op(test_operator)}(#T))"
    O = PDEOperator(APT_BilinearForm, name, [a_operator, ansatz_operator, test_operator], ConvectionAction(xdim, ncomponents, bonus_quadorder),
                    [1, 2], factor, regions, store)
    O.fixed_arguments = [a_to]
    O.fixed_arguments_ids = [a_from]
    O.transposed_assembly = transposed_assembly
    return O


def full_assemble_operator(A: FEMatrixBlock, b: FEVectorBlock, O: PDEOperator, CurrentSolution, Pattern=None, skip_preps=False):
    """the NonlinearForm branch of the operator assembly (pdeoperators.jl:1094-1140): full_assemble!(A, b, Pattern, CurrentSolution[ids];
    factor, transposed_assembly)"""
    assert O.APT == APT_NonlinearForm
    fes = CurrentSolution[O.fixed_arguments_ids[0]].FES
    if Pattern is None:
        Pattern = getattr(O, "_pattern", None)
        if Pattern is None or Pattern.FES[0] is not fes:
            Pattern = DiscreteNonlinearForm(O.operators4arguments, [fes, fes, A.FESX], O.action, name=O.name, regions=O.regions)
            O._pattern = Pattern
    full_assemble(A, b, Pattern, [CurrentSolution[O.fixed_arguments_ids[0]]], factor=O.factor, transposed_assembly=O.transposed_assembly,
                  skip_preps=skip_preps)
    return Pattern


def create_assembly_pattern(O: PDEOperator, target, CurrentSolution=None):
    """pdeoperators.jl:910-971"""
    if isinstance(target, FEMatrixBlock):
        FES = [target.FESY, target.FESX] if O.transposed_assembly else [target.FESX, target.FESY]
        if O.fixed_arguments_ids:       # 922-927: the FESpaces of the fixed arguments come first
            from .assembly import DiscreteBilinearForm
            fixedFES = [CurrentSolution[i].FES for i in O.fixed_arguments_ids]
            return DiscreteBilinearForm(O.operators4arguments, fixedFES + FES, O.action, name=O.name, regions=O.regions)
        return AssemblyPattern(O.APT, O.name, FES, O.operators4arguments, O.action, O.apply_action_to, O.regions)
    if isinstance(target, FEVectorBlock):
        if O.APT != APT_LinearForm:
            raise NotImplementedError("recasting a BilinearForm into a LinearForm needs FEB arguments (SURVEY.md 8f N4)")
        return AssemblyPattern(O.APT, O.name, [target.FES], O.operators4arguments, O.action, O.apply_action_to, O.regions)
    raise TypeError("assemble into an FEMatrixBlock or FEVectorBlock")


def assemble_operator(target, O: PDEOperator, CurrentSolution=None, Pattern=None, skip_preps=False, time=0, At=None, factor=1):
    """assemble_operator!(A::FEMatrixBlock, O; ...) / assemble_operator!(b::FEVectorBlock, O; ...)
    (pdeoperators.jl:978-1006); the flush! of the reference is implicit (results arrive as CSC)."""
    if Pattern is None:
        Pattern = getattr(O, "_pattern", None)
        if Pattern is None or Pattern.FES[0].xgrid is not (target.FESX if isinstance(target, FEMatrixBlock) else target.FES).xgrid:
            Pattern = create_assembly_pattern(O, target, CurrentSolution)
            O._pattern = Pattern
    if isinstance(target, FEMatrixBlock):
        if O.fixed_arguments_ids:      # pdeoperators.jl:986-987
            assemble(target, Pattern, [CurrentSolution[i] for i in O.fixed_arguments_ids], skip_preps=skip_preps,
                     transposed_assembly=O.transposed_assembly, factor=O.factor, fixed_arguments=O.fixed_arguments)
        elif At is not None:
            ft = O.factor if O.transpose_factor is None else O.transpose_factor
            assemble(target, Pattern, skip_preps=skip_preps, transposed_assembly=O.transposed_assembly, factor=O.factor,
                     transpose_copy=At, factor_transpose=ft)
        else:
            assemble(target, Pattern, skip_preps=skip_preps, transposed_assembly=O.transposed_assembly, factor=O.factor)
    else:
        assemble(target, Pattern, skip_preps=skip_preps, factor=O.factor * factor)
    return Pattern
