"""GPU parity of the Krylov path against the oracle's direct solve (through the C-ABI)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle.mesh import unit_cube_mesh
from oracle.mpet import MPETOracle, Coef
from tests.test_gpu_assembly import PARAMS, _engine


def _problem(n, A, theta, dt, jitter=0.2, nonsym=False):
    mesh = unit_cube_mesh(n, jitter=jitter)
    params = dict(PARAMS[A])
    if nonsym:
        params["S"] = ((0.0, 2.0), (1.0, 0.0))
    o = MPETOracle(mesh, params, dt=dt, theta=theta)
    o.momentum_markers[:] = 0
    for i in range(A):
        o.continuity_markers[i][:] = 0
    pi = np.pi
    o.u_bar = Coef(fn=lambda x, t: 0.1 * np.stack([np.cos(pi * x[:, 0]) * np.sin(pi * x[:, 1]),
                                                   np.sin(pi * x[:, 0]) * np.cos(pi * x[:, 2]),
                                                   x[:, 0] * x[:, 1]], 1) * np.sin(pi * t))
    o.p_bar = [Coef(fn=lambda x, t, i=i: (i + 1) * np.sin(pi * x[:, 0]) * np.cos(pi * x[:, 1]) * np.sin(2 * pi * t))
               for i in range(A)]
    o.f = Coef(value=(0.3, -0.2, 1.0))
    o.g = [Coef(value=0.5 + i) for i in range(A)]
    rng = np.random.default_rng(5)
    o.up_ = 0.1 * rng.standard_normal(o.space.N)
    return mesh, params, o


def _reference_solution(o):
    A = o.assemble_lhs()
    b, dofs, vals = o.rhs(o.t)
    As, bs = o.apply_bc_symmetric(A, dofs, b)
    return spla.splu(As.tocsc()).solve(bs), b, dofs, vals


def _field_errors(o, x, xref):
    sp_ = o.space
    nu = 3 * sp_.N2
    errs = [np.linalg.norm(x[:nu] - xref[:nu]) / np.linalg.norm(xref[:nu])]
    for i in range(o.A):
        d = sp_.p_dofs(i)
        errs.append(np.linalg.norm(x[d] - xref[d]) / np.linalg.norm(xref[d]))
    return errs


@pytest.mark.parametrize("pc", ["none", "jacobi", "amg"])
def test_minres_matches_direct_solve(pc):
    n, A, theta, dt = 4, 2, 1.0, 0.1
    mesh, params, o = _problem(n, A, theta, dt)
    xref, b, dofs, vals = _reference_solution(o)
    eng = _engine(mesh, params, dt, theta)
    eng.assemble_lhs()
    eng.set_dirichlet_dofs(dofs.astype(np.int32))
    eng.set_dirichlet_values(vals)
    eng.krylov_setup("minres", pc, rtol=1e-13, atol=1e-50, maxit=20000)
    if pc == "amg":
        eng.assemble_prec()
    if pc != "none":
        eng.pc_setup()
    bd = torch.as_tensor(b, device="cuda")
    x = torch.zeros_like(bd)
    info = eng.solve(bd, x)
    assert info["converged"], info
    errs = _field_errors(o, x.cpu().numpy(), xref)
    print(pc, info, errs)
    assert max(errs) < 1e-8, errs      # north_star: 1e-8 relative L2 per field
    eng.close()


def test_amg_preconditioner_is_spd_and_effective():
    n, A, theta, dt = 6, 2, 1.0, 0.1
    mesh, params, o = _problem(n, A, theta, dt)
    xref, b, dofs, vals = _reference_solution(o)
    eng = _engine(mesh, params, dt, theta)
    eng.assemble_lhs()
    eng.assemble_prec()
    eng.set_dirichlet_dofs(dofs.astype(np.int32))
    eng.set_dirichlet_values(vals)
    eng.krylov_setup("minres", "amg", rtol=1e-10, maxit=2000)
    eng.pc_setup()
    N = o.space.N
    rng = np.random.default_rng(1)
    mask = np.zeros(N, bool)
    mask[dofs] = True
    a = rng.standard_normal(N); a[mask] = 0
    c = rng.standard_normal(N); c[mask] = 0
    ad, cd = torch.as_tensor(a, device="cuda"), torch.as_tensor(c, device="cuda")
    za, zc = torch.empty_like(ad), torch.empty_like(cd)
    eng.pc_apply(ad, za)
    eng.pc_apply(cd, zc)
    sym = abs(float(za @ cd) - float(ad @ zc)) / abs(float(za @ cd))
    assert sym < 1e-10, sym
    assert float(za @ ad) > 0 and float(zc @ cd) > 0
    # one V-cycle approximates the exact block solve: the energy quotient stays within fixed bounds
    P = o.apply_bc_symmetric(o.assemble_prec(), dofs)
    exact = spla.splu(P.tocsc()).solve(a)
    q = float(za @ ad) / float(a @ exact)
    assert 0.3 < q < 1.05, q
    bd = torch.as_tensor(b, device="cuda")
    x = torch.zeros_like(bd)
    info_amg = eng.solve(bd, x)
    assert info_amg["converged"]
    # exact-block MINRES on the oracle side: the V-cycle must not cost more than ~2.5x its iterations
    its = [0]
    As, bs = o.apply_bc_symmetric(o.assemble_lhs(), dofs, b)
    lu = spla.splu(P.tocsc())
    spla.minres(As, bs, M=spla.LinearOperator(P.shape, matvec=lu.solve), rtol=1e-10, maxiter=5000,
                callback=lambda xk: its.__setitem__(0, its[0] + 1))
    print("amg its", info_amg["niter"], "exact-block its", its[0])
    assert info_amg["niter"] <= 2.5 * its[0] + 10
    eng.close()


def test_gmres_nonsymmetric_exchange():
    """The reference's own MMS test uses a non-symmetric S (test_convergence_mpetsolver.py:116)."""
    n, A, theta, dt = 4, 2, 0.5, 0.1
    mesh, params, o = _problem(n, A, theta, dt, nonsym=True)
    xref, b, dofs, vals = _reference_solution(o)
    eng = _engine(mesh, params, dt, theta)
    eng.assemble_lhs()
    eng.assemble_prec()
    eng.set_dirichlet_dofs(dofs.astype(np.int32))
    eng.set_dirichlet_values(vals)
    eng.krylov_setup("gmres", "amg", rtol=1e-13, maxit=3000, restart=60)
    eng.pc_setup()
    bd = torch.as_tensor(b, device="cuda")
    x = torch.zeros_like(bd)
    info = eng.solve(bd, x)
    errs = _field_errors(o, x.cpu().numpy(), xref)
    print(info, errs)
    assert info["converged"] and max(errs) < 1e-8
    eng.close()


def test_iteration_counts_match_oracle_restatement():
    """Same MINRES recurrence + same AMG construction on both sides -> same iteration counts
    (the reference pins no iteration counts; this pins the CUDA path to its CPU restatement)."""
    from oracle.krylov import minres, BlockAMG
    n, A, theta, dt = 6, 2, 1.0, 0.1
    mesh, params, o = _problem(n, A, theta, dt)
    Ao = o.assemble_lhs()
    b, dofs, vals = o.rhs(o.t)
    M = BlockAMG(o, dofs)
    mask = np.zeros(o.space.N, bool)
    mask[dofs] = True
    x0 = np.zeros(o.space.N)
    x0[dofs] = vals
    eng = _engine(mesh, params, dt, theta)
    eng.assemble_lhs()
    eng.assemble_prec()
    eng.set_dirichlet_dofs(dofs.astype(np.int32))
    eng.set_dirichlet_values(vals)
    for rtol in (1e-5, 1e-10):
        M = BlockAMG(o, dofs, light=rtol >= 1e-8)       # loose tolerances take the light P1-field cycles
        xo, info_o = minres(Ao, b, x0, M, mask=mask, rtol=rtol, maxit=2000)
        eng.krylov_setup("minres", "amg", rtol=rtol, maxit=2000)
        eng.pc_setup()
        bd = torch.as_tensor(b, device="cuda")
        x = torch.zeros_like(bd)
        info = eng.solve(bd, x)
        print(rtol, "gpu", info["niter"], "oracle", info_o["niter"], M.num_levels())
        assert info["converged"] and info_o["converged"]
        assert abs(info["niter"] - info_o["niter"]) <= 2
        assert np.linalg.norm(x.cpu().numpy() - xo) / np.linalg.norm(xo) < 10 * rtol
    # one preconditioner application agrees to round-off
    r = np.random.default_rng(2).standard_normal(o.space.N)
    r[mask] = 0
    z = torch.empty_like(bd)
    eng.pc_apply(torch.as_tensor(r, device="cuda"), z)
    zo = M(r)
    assert np.linalg.norm(z.cpu().numpy() - zo) / np.linalg.norm(zo) < 1e-10
    eng.close()


def test_staged_kernels_midsize():
    """n = 9: every AMG level-0 matrix is large enough for the shared-memory-staged (bulk-copy) SpMM, and
    the block SpMV runs through several chunks per CTA; compare SpMV, V-cycle and solve with the oracle."""
    from oracle.krylov import minres, BlockAMG
    n, A, theta, dt = 9, 2, 0.5, 0.1
    mesh, params, o = _problem(n, A, theta, dt)
    Ao = o.assemble_lhs()
    b, dofs, vals = o.rhs(o.t)
    eng = _engine(mesh, params, dt, theta)
    eng.assemble_lhs()
    eng.assemble_prec()
    eng.set_dirichlet_dofs(dofs.astype(np.int32))
    eng.set_dirichlet_values(vals)
    eng.krylov_setup("minres", "amg", rtol=1e-9, maxit=2000)
    eng.pc_setup()
    N = o.space.N
    rng = np.random.default_rng(7)
    xv = rng.standard_normal(N)
    xd = torch.as_tensor(xv, device="cuda")
    yd = torch.empty_like(xd)
    eng.spmv(xd, yd)
    assert np.linalg.norm(yd.cpu().numpy() - Ao @ xv) / np.linalg.norm(Ao @ xv) < 1e-13
    M = BlockAMG(o, dofs)
    mask = np.zeros(N, bool)
    mask[dofs] = True
    r = rng.standard_normal(N)
    r[mask] = 0
    z = torch.empty_like(xd)
    eng.pc_apply(torch.as_tensor(r, device="cuda"), z)
    zo = M(r)
    assert np.linalg.norm(z.cpu().numpy() - zo) / np.linalg.norm(zo) < 1e-10
    x0 = np.zeros(N)
    x0[dofs] = vals
    xo, info_o = minres(Ao, b, x0, M, mask=mask, rtol=1e-9, maxit=2000)
    x = torch.zeros_like(xd)
    info = eng.solve(torch.as_tensor(b, device="cuda"), x)
    assert info["converged"] and abs(info["niter"] - info_o["niter"]) <= 2
    assert np.linalg.norm(x.cpu().numpy() - xo) / np.linalg.norm(xo) < 1e-7
    eng.close()

