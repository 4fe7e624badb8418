"""Scale probe: timings of each phase of the hot path on one GPU (development aid)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from waterscapes_b200.workloads import make_problem, sizes
from waterscapes_b200.mpet import MPETSolver

name = sys.argv[1]; n = int(sys.argv[2]); rtol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-6
def sync(): torch.cuda.synchronize()
t0 = time.time(); problem, sp, init = make_problem(name, n); t1 = time.time()
print("host mesh+markers %.2fs" % (t1 - t0), flush=True)
solver = MPETSolver(problem, sp); sync(); t2 = time.time()
eng = solver.engine; s = eng.sizes
print("set_mesh+forms %.2fs" % (t2 - t1), s, "dev GB %.2f" % (eng.device_bytes() / 1e9), flush=True)
init(solver)
ev = lambda: torch.cuda.Event(enable_timing=True)
def timeit(fn, reps=3):
    fn(); sync(); a, b = ev(), ev(); a.record()
    for _ in range(reps): fn()
    b.record(); sync(); return a.elapsed_time(b) / reps
solver._push_params()
t_asm = timeit(eng.assemble_lhs)
print("assemble_lhs %.3f ms  (%.1f GB/s of 8*nnz)" % (t_asm, 8 * s["nnz"] / t_asm / 1e6), flush=True)
x = torch.randn(s["N"], dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
t_spmv = timeit(lambda: eng.spmv(x, y), 10)
bytes_spmv = 12 * s["nnz"] + 20 * s["N"]
print("spmv %.3f ms  %.1f GB/s" % (t_spmv, bytes_spmv / t_spmv / 1e6), flush=True)
t_rhs = timeit(lambda: eng.rhs_prev(x, y))
print("rhs_prev %.3f ms" % t_rhs, flush=True)
solver.params["krylov_rtol"] = rtol
solver.params["direct_solver"] = False
t3 = time.time()
gen = solver.solve()
for k in range(3):
    ta = time.time(); up, t = next(gen); sync(); tb = time.time()
    print("step %d: %.3f s niter %s info %s" % (k, tb - ta, solver.solver_monitor["niter"][-1], solver.solver_monitor["last"]), flush=True)
print("launches", eng.launch_count())
