"""Development aid: time the block SpMV and one preconditioner application for the pipeline configuration
selected by MPET_BLK_CFG / MPET_SPM_CFG (staged.cu), on the bench workload.  One process per configuration."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from waterscapes_b200.workloads import make_problem
from waterscapes_b200.mpet import MPETSolver

name = sys.argv[1]; n = int(sys.argv[2]); what = sys.argv[3]          # what: spmv | pc | both
problem, sp, init = make_problem(name, n)
solver = MPETSolver(problem, dict(sp, direct_solver=False))
eng = solver.engine; s = eng.sizes
init(solver)
solver._push_params()
eng.assemble_lhs()
ev = lambda: torch.cuda.Event(enable_timing=True)
def timeit(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = ev(), ev(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / reps
x = torch.randn(s["N"], dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
tag = "BLK=%s SPM=%s" % (os.environ.get("MPET_BLK_CFG", "-"), os.environ.get("MPET_SPM_CFG", "-"))
if what in ("spmv", "both"):
    t = timeit(lambda: eng.spmv(x, y), 20)
    # reference result through the generic CSR kernel
    rp, cols = eng.pattern(); vals = eng.values(0); y2 = torch.empty_like(x)
    eng.csr_spmv(rp, cols, vals, x, y2)
    err = float((y - y2).norm() / y2.norm())
    print("%s spmv(api, incl. layout conversions) %.3f ms  err %.2e" % (tag, t, err), flush=True)
if what in ("pc", "both"):
    bcs = solver.bcs[0] + solver.bcs[1]
    solver._sync_dirichlet(bcs)
    solver._ensure_prec()
    eng.krylov_setup("minres", "amg", 1e-6, 1e-50, 10000)
    eng.pc_setup()
    t = timeit(lambda: eng.pc_apply(x, y), 10)
    print("%s pc_apply(api) %.3f ms  |z| %.6e" % (tag, t, float(y.norm())), flush=True)
