"""The BASELINE.json configurations through the reference-facing API on the GPU against the oracle's sparse-LU
time loop: cfg1 at its stated size and its full 10 implicit-Euler steps (configs[0]: UnitCubeMesh(8), P2-P1-P1),
and the parameter sets / boundary conditions of cfg2 (Biot, traction-loaded), cfg3/cfg4 (four-network brain
parameters, nu = 0.4999) and cfg5 at sizes the oracle's LU finishes in seconds.  Bar: 1e-8 relative L2 per
field after every step (BASELINE.json north_star)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _field_errors(x, xo, sp_, nfields):
    nu = 3 * sp_.N2
    rel = lambda a, b: np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)
    return [rel(x[:nu], xo[:nu])] + [rel(x[sp_.p_dofs(i)], xo[sp_.p_dofs(i)]) for i in range(nfields)]


@pytest.mark.parametrize("config,n,steps", [("cfg1", 8, 10), ("cfg2", 6, 3), ("cfg3", 5, 3), ("cfg5", 6, 3)])
def test_config_matches_oracle_direct_solve(config, n, steps):
    import bench
    from waterscapes_b200.workloads import make_problem
    from waterscapes_b200.mpet import MPETSolver
    problem, sp, init = make_problem(config, n)
    T = steps * sp["dt"]
    solver = MPETSolver(problem, dict(sp, T=T, direct_solver=True))
    init(solver)
    o = bench.build_oracle_problem(config, n)
    o.T = T
    assert o.space.N == solver.engine.sizes["N"]
    if config == "cfg1":
        from waterscapes_b200.workloads import sizes
        assert solver.engine.sizes["N"] == 16197 and solver.engine.sizes["nnz"] == sizes(8, 2)["nnz"] == 1622041
    ref = [(up.copy(), t) for up, t in o.solve_direct()]
    k = 0
    J = int(problem.params["J"])
    worst = 0.0
    for up, t in solver.solve():
        xo, to = ref[k]
        assert abs(t - to) < 1e-12
        errs = _field_errors(up.vector().get_local(), xo, o.space, J)
        # a field that is (still) identically zero in the oracle has no relative error to speak of
        errs = [e for e, i in zip(errs, range(J + 1))
                if np.linalg.norm(xo[:3 * o.space.N2] if i == 0 else xo[o.space.p_dofs(i - 1)]) > 1e-30]
        assert max(errs) < 1e-8, (config, k, errs)
        worst = max(worst, max(errs))
        k += 1
    assert k == len(ref) == steps
    print(config, "n =", n, "steps =", steps, "worst field error %.2e" % worst, "iterations", solver.solver_monitor["niter"])
