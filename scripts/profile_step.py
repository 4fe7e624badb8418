"""One bench step of cfg5 inside a cudaProfilerStart/Stop bracket (for `ncu --profile-from-start off`).

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
        --log-file gpurun_out/launches.csv python scripts/profile_step.py [config] [n]

Warm-up steps (hierarchy set-up, graph capture) run outside the bracket; numbers printed under ncu are not bench values."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from waterscapes_b200.workloads import make_problem, CONFIGS
from waterscapes_b200.mpet import MPETSolver

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
n = int(sys.argv[2]) if len(sys.argv) > 2 else CONFIGS[name]["n"]
problem, sp, init = make_problem(name, n)
solver = MPETSolver(problem, dict(sp, direct_solver=False, krylov_rtol=1e-6))
init(solver)
for _ in range(2):
    solver.step(sp["dt"])
    solver.up_.assign(solver.up)
torch.cuda.synchronize()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
solver.step(sp["dt"])
torch.cuda.synchronize()
rt.cudaProfilerStop()
print("profiled step: %s iterations, %d launches so far" % (solver.solver_monitor["niter"][-1], solver.engine.launch_count()))
