"""Symmetric Dirichlet elimination -- same names and call pattern as src/mpet/mpet/bc_symmetric.py:6-22.

The reference zeroes the rows AND columns of the Dirichlet dofs in the assembled PETSc matrix
(``MatZeroRowsColumns``) and corrects the right-hand side.  On the B200 path the matrix is never touched: the
elimination is recorded on the matrix handle, the operator returns identity on the constrained rows, the initial
guess carries the boundary values and the initial residual is zeroed there, so that the Krylov solver iterates on
exactly the symmetrically eliminated system; ``mpet_get_values(which=2)`` exports the
eliminated matrix for parity checks."""
import numpy as np


def get_bc_dofs(bc):
    """bc_symmetric.py:6-8: the dofs a DirichletBC constrains."""
    return np.array(list(bc.get_boundary_values().keys()), dtype=np.intc)


def zero_rows_cols(dofs, A, b=None):
    """bc_symmetric.py:11-18 (MatZeroRowsColumns): recorded on the handle; the right-hand side is
    corrected inside the solve through the initial residual (x carries the boundary values)."""
    A.symmetric = True
    A._sym_dofs = np.asarray(dofs)


def apply_symmetric(bc, A, b=None):
    """bc_symmetric.py:20-22."""
    if bc not in A.bcs:
        A.bcs.append(bc)
    zero_rows_cols(bc.dofs(), A, b)
