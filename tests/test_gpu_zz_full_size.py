"""BASELINE.json's full size on the GPU (cfg5: 72^3 cubes, A = 3, 10.3 M dofs, 1.27 G nonzeros -- where the oracle
cannot go): size-independent properties of the assembled system, through the C-ABI.

* sizes and nnz equal the closed forms of SURVEY.md section 8; the CSR pattern is complete and sorted;
* assembling twice gives bit-identical values (deterministic gather-add, no atomics);
* A is symmetric for symmetric exchange (x.Ay == y.Ax), SpMV is linear;
* rigid translations are in the kernel of the un-constrained operator (row sums of the displacement block and of
  the divergence block cancel);
* the lumped P1 and P2 mass vectors sum to the volume of the domain."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def test_full_size_properties_cfg5():
    from waterscapes_b200.engine import Engine
    from waterscapes_b200.mpet import BoxMesh
    from waterscapes_b200.workloads import brain_params, sizes, CONFIGS
    n, J = CONFIGS["cfg5"]["n"], CONFIGS["cfg5"]["J"]
    L = 120.0
    mesh = BoxMesh((0.0, 0.0, 0.0), (L, L, L), n, n, n)
    p = brain_params(J)
    eng = Engine(0)
    S = eng.set_mesh(mesh.coordinates, mesh.cells.astype(np.int32), J)
    want = sizes(n, J)
    assert S["Nc"] == want["cells"] and S["N"] == want["dofs"] and S["nnz"] == want["nnz"]
    assert (S["nnz22"], S["nnz21"], S["nnz11"]) == (want["nnz22"], want["nnz21"], want["nnz11"])
    N, nnz = S["N"], S["nnz"]

    # ---- pattern: complete, rows sorted strictly ascending
    rowptr, cols = eng.pattern()
    assert int(rowptr[0]) == 0 and int(rowptr[-1]) == nnz
    assert bool((rowptr[1:] > rowptr[:-1]).all())
    assert int(cols.min()) == 0 and int(cols.max()) == N - 1
    d = cols[1:] - cols[:-1]
    d[rowptr[1:-1] - 1] = 1                      # differences across a row boundary do not count
    assert bool((d > 0).all())
    del d, cols, rowptr
    torch.cuda.empty_cache()

    # ---- assembly: deterministic
    eng.set_params(p["E"], p["nu"], p["alpha"], p["K"], p["S"], p["c"], CONFIGS["cfg5"]["dt"], CONFIGS["cfg5"]["theta"])
    eng.assemble_lhs()
    v1 = eng.values(0)
    eng.assemble_lhs()
    v2 = eng.values(0)
    assert torch.equal(v1, v2)
    assert bool(torch.isfinite(v1).all())
    del v1, v2
    torch.cuda.empty_cache()

    # ---- operator: linear, symmetric, translations in its kernel
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
    Ax, Ay, Az = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    eng.spmv(x, Ax)
    eng.spmv(y, Ay)
    eng.spmv(2.0 * x - 3.0 * y, Az)
    scale = float(Ax.abs().max())
    assert float((Az - (2.0 * Ax - 3.0 * Ay)).abs().max()) < 1e-12 * scale
    xAy, yAx = float(torch.dot(x, Ay)), float(torch.dot(y, Ax))
    assert abs(xAy - yAx) < 1e-10 * float(torch.linalg.norm(x)) * float(torch.linalg.norm(Ay))
    N2 = S["N2"]
    for k in range(3):
        t = torch.zeros(N, dtype=torch.float64, device="cuda")
        t[k * N2:(k + 1) * N2] = 1.0
        At = torch.empty_like(t)
        eng.spmv(t, At)
        assert float(At.abs().max()) < 1e-10 * scale, (k, float(At.abs().max()), scale)

    # ---- mass: the lumped vectors sum to the volume
    for space in (1, 2):
        w = eng.lumped(space)
        assert abs(float(w.sum()) - L ** 3) < 1e-9 * L ** 3
    eng.close()
