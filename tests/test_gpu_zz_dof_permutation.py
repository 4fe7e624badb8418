"""External dof numbering at the C-ABI (mpet_set_dof_permutation): the caller's (DOLFIN's re-ordered)
numbering for every vector and dof-index entry point, checked against the contract-numbering calls."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from tests.test_gpu_assembly import _engine
from tests.test_gpu_solve import _problem, _reference_solution, _field_errors


def test_external_dof_permutation():
    """mpet_set_dof_permutation: with a random renumbering of the dofs at the boundary, every vector
    and dof-index entry point gives the contract-numbering result re-indexed -- bit for bit for the products, to the
    solver tolerance for the solve."""
    n, A, theta, dt = 4, 2, 1.0, 0.1
    mesh, params, o = _problem(n, A, theta, dt)
    xref, b, dofs, vals = _reference_solution(o)
    N = o.space.N
    rng = np.random.default_rng(11)
    perm = rng.permutation(N).astype(np.int32)            # perm[c] = caller's index of contract dof c
    inv = np.empty(N, dtype=np.int64)
    inv[perm] = np.arange(N)

    def ext(v):                                           # contract-numbered host vector -> caller's numbering
        out = np.empty_like(v)
        out[perm] = v
        return out

    eng = _engine(mesh, params, dt, theta)
    eng.assemble_lhs()
    eng.assemble_prec()
    x0 = rng.standard_normal(N)
    xd = torch.as_tensor(x0, device="cuda")
    y_con, r_con = torch.empty_like(xd), torch.empty_like(xd)
    eng.spmv(xd, y_con)
    eng.rhs_prev(xd, r_con)
    # --- caller's numbering
    eng.set_dof_permutation(perm)
    xe = torch.as_tensor(ext(x0), device="cuda")
    y_ext, r_ext = torch.empty_like(xe), torch.empty_like(xe)
    eng.spmv(xe, y_ext)
    eng.rhs_prev(xe, r_ext)
    assert np.array_equal(y_ext.cpu().numpy(), ext(y_con.cpu().numpy()))
    assert np.array_equal(r_ext.cpu().numpy(), ext(r_con.cpu().numpy()))
    eng.set_dirichlet_dofs(perm[dofs].astype(np.int32))   # caller's indices of the constrained dofs, same order
    eng.set_dirichlet_values(vals)
    eng.krylov_setup("minres", "amg", rtol=1e-13, atol=1e-50, maxit=20000)
    eng.pc_setup()
    bd = torch.as_tensor(ext(b), device="cuda")
    x = torch.zeros_like(bd)
    info = eng.solve(bd, x)
    assert info["converged"], info
    errs = _field_errors(o, x.cpu().numpy()[perm], xref)  # back to the contract numbering
    print("permuted solve", info["niter"], errs)
    assert max(errs) < 1e-8, errs
    # the matrix exports stay in the contract numbering; a non-bijection is refused; NULL restores the contract
    with pytest.raises(Exception):
        eng.set_dof_permutation(np.zeros(N, dtype=np.int32))
    eng.set_dof_permutation(None)
    y2 = torch.empty_like(xd)
    eng.spmv(xd, y2)
    assert np.array_equal(y2.cpu().numpy(), y_con.cpu().numpy())
    eng.close()
