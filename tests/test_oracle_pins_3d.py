"""CPU: oracle-independent pins of the oracle's 3-D restatement (the reference's own tests
are 2-D, the product is 3-D).  The 3-D twins of the reference's tests, with the reference's assertions:

* src/mpet/test/test_convergence_mpetsolver.py:103-245 and test_convergence_totalpressuresolver.py:103-253 with the
  3-D extension of their manufactured solutions (tests/mms3d.py: sympy-derived f, g, total stress; errors integrated
  against the exact solution, no oracle code in the error norms) on UnitCubeMesh(4), (8): the reference's rate
  thresholds already hold on this coarse pair (measured 2.98 / 1.84 / 2.7 / 1.49 for the standard formulation);
* src/mpet/test/test_donut.py:16-71 on a jittered 3-D box: traction s = t*n on the whole boundary, six rigid-motion
  Lagrange multipliers (rm_basis_L2.py:10-74), exact solution p = -t.

The same checks run on the GPU path in tests/test_gpu_convergence.py / test_gpu_known_answer.py (n up to 32)."""
import numpy as np
import pytest

from oracle.mesh import unit_cube_mesh
from oracle.mpet import MPETOracle, MPETTotalPressureOracle, Coef
from tests import mms3d


def _mms_run(n, M, theta, total_pressure):
    T = 1.0
    if total_pressure:      # test_convergence_totalpressuresolver.py:110-118
        params = dict(J=2, c=(1.0, 1.0), alpha=(1.0, 1.0), K=(1.0, 1.0), S=((1.0, 1.0), (1.0, 1.0)), E=1.0, nu=0.35)
    else:                   # test_convergence_mpetsolver.py:110-118
        params = dict(J=2, c=(0.3, 0.4), alpha=(0.4, 0.6), K=(0.2, 0.3), S=((0.0, 2.0), (1.0, 0.0)), E=520.0, nu=0.47)
    ex = mms3d.exact_solutions(params)
    mesh = unit_cube_mesh(n)
    o = (MPETTotalPressureOracle if total_pressure else MPETOracle)(mesh, params, dt=T / M, theta=theta, T=T)
    o.f = Coef(fn=ex["f"], degree=3)
    o.g = [Coef(fn=ex["g"][i], degree=3) for i in range(2)]
    o.u_bar = Coef(fn=ex["u"], degree=3)
    o.s = Coef(fn=ex["sigma"], degree=4)
    o.s_times_normal = True
    o.p_bar = [Coef(fn=ex["p"][i], degree=3) for i in range(2)]
    xm = mesh.coords[o.facets["vertices"]]
    o.momentum_markers[:] = 0
    o.momentum_markers[np.all(np.abs(xm[:, :, 0] - 1.0) < 3e-16, axis=1)] = 1      # near(x[0], 1.0): traction
    for i in range(2):
        o.continuity_markers[i][:] = 0
    sp_ = o.space
    u0 = ex["u"](sp_.node2_coords(), 0.0)
    for k in range(3):
        o.up_[sp_.u_dofs(k)] = u0[:, k]
    off = 0
    if total_pressure:
        o.up_[sp_.p_dofs(0)] = ex["p_total"](mesh.coords, 0.0)
        off = 1
    for i in range(2):
        o.up_[sp_.p_dofs(i + off)] = ex["p"][i](mesh.coords, 0.0)
    for up, t in o.solve_direct():
        pass
    u = np.stack([up[sp_.u_dofs(k)] for k in range(3)], axis=1)
    ps = [up[sp_.p_dofs(i + off)] for i in range(2)]
    ev = sp_.edge_vertices
    keys = ev[:, 0] * sp_.Nv + ev[:, 1]

    def edge_index(lo, hi):
        return np.searchsorted(keys, np.asarray(lo, dtype=np.int64) * sp_.Nv + np.asarray(hi, dtype=np.int64))

    err = mms3d.error_norms(mesh.coords, mesh.cells, edge_index, sp_.Nv, u, ps, ex, float(t))
    return err, mms3d.hmin(mesh.coords, mesh.cells)


@pytest.mark.parametrize("total_pressure", [False, True])
def test_mms_convergence_rates_3d_oracle(total_pressure):
    res = [_mms_run(n, m, 0.5, total_pressure) for n, m in ((4, 2), (8, 4))]
    hs = [r[1] for r in res]
    u_L2 = mms3d.rates([r[0]["u_L2"] for r in res], hs)[-1]
    u_H1 = mms3d.rates([r[0]["u_H1"] for r in res], hs)[-1]
    p_L2 = [mms3d.rates([r[0]["p_L2"][i] for r in res], hs)[-1] for i in range(2)]
    p_H1 = [mms3d.rates([r[0]["p_H1"][i] for r in res], hs)[-1] for i in range(2)]
    print("total_pressure" if total_pressure else "standard", u_L2, u_H1, p_L2, p_H1)
    assert u_L2 > 1.70 and u_H1 > 1.70
    if total_pressure:      # test_convergence_totalpressuresolver.py:237-242
        assert p_L2[0] > 1.85 and p_L2[1] > 1.87
    else:                   # test_convergence_mpetsolver.py:234-239
        assert p_L2[0] > 1.70 and p_L2[1] > 1.70
    assert p_H1[0] > 0.95 and p_H1[1] > 0.95


def test_constant_pressure_on_a_3d_box_with_rigid_motions():
    """test_donut.py:16-71 with its own data in 3-D: u determined up to the SIX rigid motions, removed by Lagrange
    multipliers; p = -t exactly (P1 reproduces constants), u = 0."""
    mesh = unit_cube_mesh(3, jitter=0.2)
    mesh.coords[:] = mesh.coords * np.array([100.0, 80.0, 60.0])
    params = dict(J=1, c=(0.0,), alpha=(1.0,), K=(1.0e-2,), S=((0.0,),), E=500, nu=0.35)
    o = MPETOracle(mesh, params, dt=0.1, theta=1.0, T=0.2, u_has_nullspace=True)
    o.s = Coef(value=lambda t: t)
    o.s_times_normal = True
    o.momentum_markers[:] = 1
    o.continuity_markers[0][:] = 0
    o.p_bar = [Coef(value=lambda t: -t)]
    for up, t in o.solve_direct():
        pass
    assert abs(t - 0.2) < 1e-12 and len(up) == o.space.nfe + 6
    u, p = o.split(up)
    assert np.max(np.abs(p[0] + 0.2)) < 1e-8                                   # test_donut.py:70 at every vertex
    vol = mesh.cell_volumes()
    M = (np.ones((4, 4)) + np.eye(4)) / 20.0
    pl = p[0][mesh.cells]
    l2 = np.sqrt(np.einsum("c,cm,mn,cn->", vol, pl, M, pl))
    assert abs(l2 / np.sqrt(vol.sum()) - 0.2) < 1e-10                          # test_donut.py:71
    assert np.max(np.abs(u)) < 1e-8 * 100.0
