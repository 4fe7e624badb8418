"""Symmetric Dirichlet elimination helpers -- same names as src/mpet/mpet/bc_symmetric.py:6-22."""
from .la import get_bc_dofs, zero_rows_cols, apply_symmetric  # noqa: F401
