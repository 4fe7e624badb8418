"""Drop-in import surface of the reference's ``from mpet import *`` (src/mpet/mpet/__init__.py:10-17)
for the assemble + solve hot path, backed by libmpet_b200 (sm_100a).  Importing this package loads the
CUDA library and raises if it is missing: there is no CPU fallback."""
from .. import _lib as _lib
_lib.load()

from .dolfin_shim import (Mesh, BoxMesh, UnitCubeMesh, Constant, Expression, FacetNormal, MeshFunction,  # noqa
                          SubDomain, CompiledSubDomain, Function, FunctionSpace, DirichletBC, Parameters,
                          parameters, interpolate, assign, info, warning, INVALID, CellFunction, near, DOLFIN_EPS)
from .la import assemble, LUSolver, PETScKrylovSolver, Matrix, Form  # noqa
from .mpetproblem import MPETProblem, convert_to_E_nu, convert_to_mu_lmbda, elastic_stress  # noqa
from .mpetsolver import MPETSolver, DIRICHLET_MARKER, NEUMANN_MARKER, ROBIN_MARKER  # noqa
from .mpettotalpressuresolver import MPETTotalPressureSolver  # noqa
from .bc_symmetric import get_bc_dofs, zero_rows_cols, apply_symmetric  # noqa
from .rm_basis_L2 import rigid_motions  # noqa  (src/mpet/mpet/__init__.py:15)
from .hdf5 import HDF5File  # noqa
