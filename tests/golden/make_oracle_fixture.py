"""Writes tests/golden/oracle_cube.npz: the oracle's own outputs on two tiny 3-D cases, both formulations.

Not a reference-derived fixture (DOLFIN cannot run here): it freezes the oracle, so that a later
edit of oracle/ that changes the pattern, the assembled values, the right-hand side or the solution is caught by
tests/test_oracle_golden.py, and gives the GPU parity tests a second, committed anchor.

    python tests/golden/make_oracle_fixture.py
"""
import os
import sys

import numpy as np
import scipy.sparse.linalg as spla

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from oracle.mesh import unit_cube_mesh                               # noqa: E402
from oracle.mpet import MPETOracle, MPETTotalPressureOracle, Coef    # noqa: E402

PARAMS = dict(J=2, E=2.2, nu=0.4545454545, alpha=(0.5, 0.5), c=(1.0, 1.0), K=(1.0, 1.0), S=((0, 1.0), (1.0, 0)))


def case(cls, n, jitter, theta=0.5, dt=0.1):
    mesh = unit_cube_mesh(n, jitter=jitter)
    o = cls(mesh, PARAMS, dt=dt, theta=theta)
    o.momentum_markers[:] = 0
    for i in range(2):
        o.continuity_markers[i][:] = 0
    pi = np.pi
    o.u_bar = Coef(fn=lambda x, t: 0.1 * np.stack([np.cos(pi * x[:, 0]), x[:, 1] * x[:, 2], np.sin(pi * x[:, 2])], 1) * (1 + t))
    o.p_bar = [Coef(fn=lambda x, t, i=i: (i + 1) * x[:, 0] * (1 + t)) for i in range(2)]
    o.f = Coef(value=(0.3, -0.2, 1.0))
    o.g = [Coef(value=0.5), Coef(value=1.5)]
    o.up_ = 0.01 * np.sin(np.arange(o.space.N) * 0.7)
    A = o.on_pattern(o.assemble_lhs())
    P = o.on_pattern(o.assemble_prec())
    b, dofs, vals = o.rhs(0.0)
    As, bs = o.apply_bc_symmetric(A, dofs, b)
    x = spla.splu(As.tocsc()).solve(bs)
    return o, A, P, b, dofs, x


def main():
    out = {}
    for tag, cls in (("std", MPETOracle), ("tp", MPETTotalPressureOracle)):
        o, A, P, b, dofs, x = case(cls, 1, 0.0)
        out[tag + "1_indptr"] = A.indptr.astype(np.int64)
        out[tag + "1_indices"] = A.indices.astype(np.int32)
        out[tag + "1_A"] = A.data
        out[tag + "1_P"] = P.data
        out[tag + "1_b"] = b
        out[tag + "1_dofs"] = dofs.astype(np.int64)
        out[tag + "1_x"] = x
        out[tag + "1_cell_dofs"] = o.space.cell_dofs.astype(np.int32)
        o, A, P, b, dofs, x = case(cls, 2, 0.2)
        out[tag + "2_sums"] = np.array([A.nnz, np.abs(A.data).sum(), np.linalg.norm(A.data), np.abs(P.data).sum(),
                                        np.linalg.norm(b), np.linalg.norm(x), float(dofs.sum())])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_cube.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
