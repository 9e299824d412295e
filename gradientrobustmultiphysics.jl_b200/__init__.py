"""grmp_b200 -- B200-native assembly engine behind the GradientRobustMultiPhysics.jl
assembly API (BilinearForm / LinearForm AssemblyPatterns).

Layout:
  csrc/          CUDA kernels + the C-ABI library libgrmp_cuda.so (include/grmp.h)
  grid.py        ExtendableGrid components the path reads (input producer stand-in)
  quadrature.py  QuadratureRule tables           (src/quadrature.jl)
  fedefs.py      FETypes + reference bases       (src/fedefs/*.jl)
  fespace.py     FESpace / CellDofs / FEVector / FEMatrix
  assembly.py    AssemblyPattern, assemble!      (src/assemblypatterns/*.jl)
  operators.py   LaplaceOperator ... assemble_operator!   (src/pdeoperators.jl)
  boundarydata.py  Dirichlet boundary data: interpolated / homogeneous / best-approximation (src/boundarydata.jl)
"""
from .grid import (ExtendableGrid, grid_unitsquare, grid_unitcube, reference_domain, uniform_refine,
                   perturb_interior_nodes)
from .quadrature import QuadratureRule
from .fedefs import H1P1, H1P2, H1Pk, H1BR, HDIVRT0, HDIVBDM1, L2P0, reference_tables
from .fespace import FESpace, FEVector, FEMatrix, FEMatrixBlock, FEVectorBlock
from .assembly import (Identity, NormalFlux, fdotn_action, Gradient, SymmetricGradient, Divergence, ReconstructionIdentity, NoAction, HookeAction, ConvectionAction, NewtonConvectionAction, DiscreteNonlinearForm, full_assemble, Action,
                       DataFunction, fdot_action, AssemblyPattern, DiscreteBilinearForm, DiscreteSymmetricBilinearForm,
                       DiscreteLumpedBilinearForm, DiscreteLinearForm, prepare_assembly, assemble, assemble_csc, blf_set_path,
                       blf_stats, quadrature_order, device_grid, device_space, addblock_matmul, residual, apply_penalties,
                       device_csc, fetch_values, ItemIntegrator, L2NormIntegrator, L2ErrorIntegrator, evaluate, evaluate_itemwise)
from .operators import (PDEOperator, LaplaceOperator, ReactionOperator, LagrangeMultiplier, ConvectionOperator, full_assemble_operator, HookStiffnessOperator2D,
                        HookStiffnessOperator3D, BilinearForm, LinearForm, create_assembly_pattern, assemble_operator)
from .boundarydata import (BoundaryData, boundarydata, HomogeneousDirichletBoundary, InterpolateDirichletBoundary,
                           BestapproxDirichletBoundary)
from . import _lib, assembly, partition
