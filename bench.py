#!/usr/bin/env python
"""Benchmark of the MPET timestep hot path (assemble + Krylov solve) on B200.

    python bench.py --gpus N --steps K --warmup W          # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus N ...          # CPU restatement of the reference path

One "step" = the body of the reference's MPETSolver.step (src/mpet/mpet/mpetsolver.py:317-379):
assemble A (+ Robin), Dirichlet conditions, assemble b (L, L1[i], L0), solve to the reference
tolerance.  Metric: whole-job DOF/s (unknowns of the block system advanced one timestep per second).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mpet_timestep_dof_per_s"
UNIT = "DOF/s"
DEFAULT_CONFIG = "cfg5"          # A=3, n=72 per GPU: ~10.3 M dofs per GPU (BASELINE.json configs[4])
CPU_SAMPLE_N = 32                # bounded CPU sample of the same workload family (966 k dofs; ~5 s per step on 32 cores)
CPU_TREND_NS = (12, 20)          # two more CPU sizes: the DOF/s trend towards the bench size is evidence, not hope
CONFIG_N = {"cfg1": 8, "cfg2": 34, "cfg3": 28, "cfg4": 71, "cfg5": 72}      # waterscapes_b200/workloads.py CONFIGS


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=DEFAULT_CONFIG)
    ap.add_argument("--mesh-n", dest="n", type=int, default=None, help="override the mesh resolution of the config")
    ap.add_argument("--rtol", type=float, default=1e-6,
                    help="Krylov tolerance (reference authors' intent: sandbox/2D_1net_totalpressure.py:168-170)")
    ap.add_argument("--maxit", type=int, default=10000)
    ap.add_argument("--cpu-n", type=int, default=CPU_SAMPLE_N)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg1-cfg4 table of the N=1 line")
    ap.add_argument("--partition", default="brick", choices=["brick", "slab"])
    ap.add_argument("--formulation", default="standard", choices=["standard", "total-pressure"],
                    help="standard = MPETSolver (the headline); total-pressure = MPETTotalPressureSolver on the same "
                         "workload (SURVEY.md 8f.1), reported as extra information")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.gpu = gpu_index
        self.path = "/tmp/bench_clocks_%d_%d.csv" % (os.getpid(), gpu_index)

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            busy = [v for v in sm if v > 0.5 * max(sm)] or sm
            out = {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def workload_string(config, n, rtol):
    """The workload both arms name (the reference arm times a bounded sample of it, `cpu_baseline.sample`).  Sizes from
    the closed forms of SURVEY.md section 8 so that the CPU arm needs no engine."""
    J = {"cfg1": 2, "cfg2": 1, "cfg3": 4, "cfg4": 4, "cfg5": 3}[config]
    dofs = 3 * (2 * n + 1) ** 3 + J * (n + 1) ** 3
    return ("%s: [P2]^3x[P1]^%d MPET on BoxMesh(%d^3) per GPU, %d cells, %d dofs per GPU; step = assemble A + b, "
            "Dirichlet, MINRES+block-AMG to rtol %g (PETSc default test: relative to the preconditioned norm of b)"
            % (config, J, n, 6 * n ** 3, dofs, rtol))


def measured_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per SpMV launch from the committed ncu --set full capture, if present."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "spmv_traffic.json")))
        return d
    except Exception:
        return None


# ------------------------------------------------------------------------------------------ CPU arm
def build_oracle_problem(config, n):
    """The workload of waterscapes_b200/workloads.py restated on the oracle's classes (no product code)."""
    import numpy as np
    from oracle.mesh import unit_cube_mesh, SimplexMesh
    from oracle.mpet import MPETOracle, Coef
    mm = 133.32
    if config == "cfg2":
        mesh = unit_cube_mesh(n)
        params = dict(J=1, E=500.0, nu=0.49, alpha=(1.0,), c=(1.0e-2,), K=(1.0e-5,), S=((0.0,),))
        o = MPETOracle(mesh, params, dt=0.05, theta=1.0)
        xm = mesh.coords[o.facets["vertices"]]
        o.momentum_markers[:] = 1
        o.momentum_markers[np.all(np.abs(xm[:, :, 2]) < 3e-16, axis=1)] = 0
        o.s = Coef(value=lambda t: 133.322 * 0.15 * np.sin(2 * np.pi * t))
        o.s_times_normal = True
        o.continuity_markers[0][:] = 1
        o.continuity_markers[0][np.all(np.abs(xm[:, :, 2] - 1.0) < 3e-16, axis=1)] = 0
        return o
    if config == "cfg1":
        mesh = unit_cube_mesh(n)
        mu, lm = 1.0, 10.0
        E, nu = mu * (3 * lm + 2 * mu) / (lm + mu), lm / (2 * (lm + mu))
        params = dict(J=2, E=E, nu=nu, alpha=(0.5, 0.5), c=(1.0, 1.0), K=(1.0, 1.0), S=((0.0, 1.0), (1.0, 0.0)))
        o = MPETOracle(mesh, params, dt=0.1, theta=1.0)
        pi = np.pi
        o.u_bar = Coef(fn=lambda x, t: 0.1 * np.stack(
            [np.cos(pi * x[:, 0]) * np.sin(pi * x[:, 1]) * np.sin(pi * x[:, 2]),
             np.sin(pi * x[:, 0]) * np.cos(pi * x[:, 1]) * np.sin(pi * x[:, 2]),
             np.sin(pi * x[:, 0]) * np.sin(pi * x[:, 1]) * np.cos(pi * x[:, 2])], 1) * np.sin(pi * t))
        o.p_bar = [Coef(fn=lambda x, t, i=i: (i + 1) * np.sin(pi * x[:, 0]) * np.cos(pi * x[:, 1])
                        * np.sin(pi * x[:, 2]) * np.sin(2 * pi * t)) for i in range(2)]
        from waterscapes_b200.mms import cfg1_exact, standard_sources      # input data only (sympy), no engine
        mms = standard_sources(params, *cfg1_exact())
        o.f = Coef(fn=mms["f"], degree=2)
        o.g = [Coef(fn=mms["g"][i], degree=1) for i in range(2)]
        o.momentum_markers[:] = 0
        for i in range(2):
            o.continuity_markers[i][:] = 0
        x2 = o.space.node2_coords()
        u0 = o.u_bar.at_points(x2, 0.0)
        for k in range(3):
            o.up_[o.space.u_dofs(k)] = u0[:, k]
        return o
    J = {"cfg3": 4, "cfg4": 4, "cfg5": 3}[config]
    base = unit_cube_mesh(n)
    mesh = SimplexMesh(base.coords * 120.0, base.cells)
    c = (3.9e-4, 2.9e-4, 1.5e-5, 2.9e-4)
    alpha = (0.49, 0.25, 0.01, 0.25)
    kappa = (1.4e-14, 1.e-10, 1.e-10, 1.e-10)
    eta = (8.9e-4, 2.67e-3, 2.67e-3, 2.67e-3)
    K = [kappa[i] / eta[i] * 1.e6 for i in range(4)]
    s = 1.0e-6
    S = ((0.0, 0.0, s, s), (0.0, 0.0, 0.0, s), (s, 0.0, 0.0, s), (s, s, s, 0.0))
    params = dict(J=J, alpha=alpha[:J], K=K[:J], S=tuple(r[:J] for r in S[:J]), c=c[:J], nu=0.4999, E=1500)
    o = MPETOracle(mesh, params, dt=0.0125, theta=0.5)
    pbar = [Coef(value=lambda t: mm * (3.0 + 2 * np.sin(2 * np.pi * t))),
            Coef(value=lambda t: mm * (70.0 + 10.0 * np.sin(2.0 * np.pi * t))),
            Coef(value=mm * 6.0), Coef(value=mm * (6.0 + 70) / 2)]
    o.p_bar = pbar[:J]
    o.momentum_markers[:] = 0
    for i in range(J):
        o.continuity_markers[i][:] = 1 if i == 3 else 0
        o.up_[o.space.p_dofs(i)] = float(pbar[i].const(0.0))
    return o


def run_cpu(config, n, rtol, steps, warmup, maxit, keep=False):
    """The oracle's step on the host, all host cores: OpenMP cell loop for assemble(a), numpy for the (small)
    right-hand side, MINRES + block AMG with every sparse product in an OpenMP row loop (oracle/omp.py,
    oracle/csrc/cpu_kernels.c).  Returns a dict: sec (per timed step), dofs, iters, threads, setup_s (AMG
    hierarchy, outside the step on both arms) and -- keep=True -- the oracle, its last matrix and state."""
    import numpy as np
    from oracle.krylov import minres, BlockAMG
    from oracle import omp
    from threadpoolctl import threadpool_limits
    threadpool_limits(1, user_api="blas")       # numpy's BLAS pool would fight the OpenMP row loops for the cores
    threads = omp.use_all_cores()
    o = build_oracle_problem(config, n)
    dofs, _ = o.dirichlet(o.t)
    t0 = time.perf_counter()
    M = omp.parallelise(BlockAMG(o, dofs, light=rtol >= 1e-8))      # hierarchy set-up is outside the step on both arms
    setup_s = time.perf_counter() - t0
    B = o.assemble_prev_operator()
    mask = np.zeros(o.space.N, dtype=bool)
    mask[dofs] = True
    x = o.up_.copy()
    no_robin = all(not np.any(m == 2) for m in o.continuity_markers)
    asm = omp.LhsAssembler(o) if no_robin else o.assemble_lhs      # pattern / tables once, like DOLFIN's first assemble
    times, iters = [], []
    A = None
    for k in range(warmup + steps):
        t0 = time.perf_counter()
        A = asm()                               # re-assembled every step like MPETSolver.step
        b, _, vals = o.rhs(o.t, B)
        x0 = x.copy()
        x0[dofs] = vals
        x, info = minres(omp.OmpCsr(A), b, x0, M, mask=mask, rtol=rtol, maxit=maxit)
        t1 = time.perf_counter()
        o.t += o.dt
        o.up_ = x.copy()
        if k >= warmup:
            times.append(t1 - t0)
            iters.append(info["niter"])
    out = dict(sec=sum(times) / len(times), dofs=o.space.N, iters=iters, threads=threads, setup_s=setup_s)
    if keep:
        out.update(oracle=o, A=A, x=x)
    return out


# ------------------------------------------------------------------------------------------ GPU helpers
def gpu_time_config(config, n, rtol, maxit, steps, warmup, device, formulation="standard", keep=False):
    """`steps` timed steps (after `warmup`) of one config on ONE GPU through MPETSolver.step; CUDA events."""
    import torch
    from waterscapes_b200.workloads import make_problem
    from waterscapes_b200.mpet import MPETSolver, MPETTotalPressureSolver
    problem, sp, init = make_problem(config, n)
    sp = dict(sp, direct_solver=False, krylov_rtol=rtol, krylov_maxit=maxit)
    tp = formulation == "total-pressure"
    solver = (MPETTotalPressureSolver if tp else MPETSolver)(problem, sp, device=device)
    if tp:
        problem.time.assign(0.0)
        for i in range(int(problem.params["J"])):
            solver.up_.set_sub(solver._net_sub(i), problem.p_bar[i])
    else:
        init(solver)
    dt = sp["dt"]
    for _ in range(warmup):
        solver.step(dt)
        solver.up_.assign(solver.up)
    torch.cuda.synchronize()
    solver.engine.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        solver.step(dt)
        solver.up_.assign(solver.up)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    S = solver.engine.sizes
    out = dict(config=config, n=n, dofs=S["N"], nnz=S["nnz"], networks=S["A"], ms_per_step=ms,
               dof_per_s=S["N"] / (ms / 1e3), krylov_iterations=solver.solver_monitor["niter"][-steps:],
               method=solver.solver_monitor.get("method"), launches_per_step=solver.engine.launch_count() / steps,
               amg_setup_s=solver.solver_monitor.get("pc_setup_s"))
    if keep:
        out["solver"] = solver
    else:
        solver.engine.close()
    return out


def true_residual(solver, partition, dist, world):
    """||b - A x||_2 / ||b||_2 of the step just solved, over free (non-Dirichlet) owned rows, b after the symmetric
    Dirichlet elimination (the system MINRES iterates on); all-reduced over ranks."""
    import numpy as np
    import torch
    eng = solver.engine
    x, b = solver.up.x, solver._last_b.t
    bcs = solver.bcs[0] + solver.bcs[1]
    D = np.concatenate([bc.dofs() for bc in bcs]) if bcs else np.zeros(0, dtype=np.int64)
    Dt = torch.as_tensor(D, device=x.device, dtype=torch.int64)
    free = torch.ones_like(x, dtype=torch.bool)
    free[Dt] = False
    if partition is not None:
        free &= torch.as_tensor(partition.owned_dofs.astype(bool), device=x.device)
    y = torch.empty_like(x)
    eng.spmv(x, y)
    r = (b - y)[free]
    xd = torch.zeros_like(x)
    xd[Dt] = x[Dt]
    eng.spmv(xd, y)
    be = (b - y)[free]
    t = torch.stack([torch.sum(r * r), torch.sum(be * be)])
    if world > 1:
        dist.all_reduce(t)
    return float(torch.sqrt(t[0] / t[1])), float(torch.sqrt(t[0])), float(torch.sqrt(t[1]))


def _field_errors(x, ref, blocks):
    import numpy as np
    errs, lo = [], 0
    for blk in blocks:
        d, r = x[lo:lo + blk] - ref[lo:lo + blk], ref[lo:lo + blk]
        errs.append(float(np.linalg.norm(d) / max(np.linalg.norm(r), 1e-300)))
        lo += blk
    return errs


def parity_single(config, device):
    """N = 1: two steps of a small instance of the bench workload on the GPU (tight 'direct' preset) against the
    oracle's sparse-LU time loop, per-field relative L2."""
    from waterscapes_b200.workloads import make_problem
    from waterscapes_b200.mpet import MPETSolver
    n, steps = 6, 2
    problem, sp, init = make_problem(config, n)
    solver = MPETSolver(problem, dict(sp, T=steps * sp["dt"], direct_solver=True), device=device)
    init(solver)
    o = build_oracle_problem(config, n)
    o.T = steps * sp["dt"]
    ref = [up.copy() for up, t in o.solve_direct()]
    worst = 0.0
    for k, (up, t) in enumerate(solver.solve()):
        sp_ = o.space
        errs = _field_errors(up.vector().get_local(), ref[k], [3 * sp_.N2] + [sp_.Nv] * int(problem.params["J"]))
        worst = max(worst, max(errs))
    its = list(solver.solver_monitor["niter"])
    solver.engine.close()
    return {"against": "oracle sparse LU (oracle/mpet.py solve_direct)", "instance": "%s at n=%d, %d steps, rtol 1e-12" % (config, n, steps),
            "max_field_rel_err": worst, "iterations": its}


def parity_distributed(config, world, rank, local_rank, dist):
    """N > 1: two steps of a small instance of the bench workload, partitioned exactly like the bench (bricks, one
    ghost layer, peer/NCCL halos, distributed V-cycles), against the SAME instance solved on one GPU by rank 0;
    per-field relative L2 over the owned dofs of all ranks."""
    import numpy as np
    from waterscapes_b200.parallel import box_brick, brick_grid, Partition, node_global_keys
    from waterscapes_b200.workloads import make_problem
    from waterscapes_b200.mpet import MPETSolver, BoxMesh
    grid = brick_grid(world)
    dims = tuple(4 * g for g in grid)
    L = 1.0 if config in ("cfg1", "cfg2") else 120.0
    p1 = tuple(L * d / 8.0 for d in dims)
    steps = 2
    local = box_brick((0.0, 0.0, 0.0), p1, dims[0], dims[1], dims[2], rank, world, grid=grid)
    problem, sp, init = make_problem(config, dims[0], mesh=local)
    sp = dict(sp, T=steps * sp["dt"], direct_solver=True)
    solver = MPETSolver(problem, sp, device=local_rank, partition=Partition(rank, world))
    init(solver)
    for up, t in solver.solve():
        pass
    xl = up.vector().get_local()
    keys = node_global_keys(solver.VQ, local)
    owned_nodes = solver.partition.owned_nodes
    ls = solver.VQ
    J = ls.J
    payload = dict(keys=keys[owned_nodes], nv_owned=int(owned_nodes[:ls.Nv].sum()),
                   u=[xl[k * ls.N2:(k + 1) * ls.N2][owned_nodes] for k in range(3)],
                   p=[xl[3 * ls.N2 + i * ls.Nv: 3 * ls.N2 + (i + 1) * ls.Nv][owned_nodes[:ls.Nv]] for i in range(J)],
                   its=list(solver.solver_monitor["niter"]))
    solver.engine.close()
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    out = None
    if rank == 0:
        gmesh = BoxMesh((0.0, 0.0, 0.0), p1, *dims)
        gproblem, gsp, ginit = make_problem(config, dims[0], mesh=gmesh)
        gsolver = MPETSolver(gproblem, dict(gsp, T=steps * gsp["dt"], direct_solver=True), device=local_rank)
        ginit(gsolver)
        for gup, gt in gsolver.solve():
            pass
        xg = gup.vector().get_local()
        gs = gsolver.VQ
        ev = gs.edge_vertices()
        nvg = gmesh.num_vertices()
        gkeys = np.concatenate([np.arange(nvg), nvg + ev[:, 0] * nvg + ev[:, 1]])
        order = np.argsort(gkeys)
        num = np.zeros(1 + J)
        den = np.zeros(1 + J)
        count = 0
        for pl in gathered:
            gnode = order[np.searchsorted(gkeys[order], pl["keys"])]
            count += gnode.size
            for k in range(3):
                ref = xg[k * gs.N2 + gnode]
                num[0] += np.sum((pl["u"][k] - ref) ** 2)
                den[0] += np.sum(ref ** 2)
            gv = gnode[:pl["nv_owned"]]
            for i in range(J):
                ref = xg[3 * gs.N2 + i * gs.Nv + gv]
                num[1 + i] += np.sum((pl["p"][i] - ref) ** 2)
                den[1 + i] += np.sum(ref ** 2)
        assert count == gs.N2, "owned nodes of all ranks must tile the mesh"
        errs = [float(np.sqrt(a / max(b, 1e-300))) for a, b in zip(num, den)]
        out = {"against": "the same instance solved on one GPU (rank 0)",
               "instance": "%s on BoxMesh(%d,%d,%d) cut into %dx%dx%d bricks, %d steps, rtol 1e-12" % ((config,) + dims + grid + (steps,)),
               "field_rel_err": errs, "max_field_rel_err": max(errs), "iterations_dist": gathered[0]["its"],
               "iterations_single": list(gsolver.solver_monitor["niter"])}
        gsolver.engine.close()
    dist.barrier()
    return out


def matched_block(config, n, rtol, maxit, device):
    """Like-for-like GPU / CPU comparison at the CPU arm's size: same workload, same n, same steps, same tolerance,
    with the assembled-matrix Frobenius difference and the per-field L2 difference of the two final states
    (BASELINE.md 3.3).  Also the CPU DOF/s at two smaller sizes (trend towards the bench size)."""
    import numpy as np
    warm, steps = 1, 2
    g = gpu_time_config(config, n, rtol, maxit, steps, warm, device, keep=True)
    solver = g.pop("solver")
    c = run_cpu(config, n, rtol, steps, warm, maxit, keep=True)
    o = c["oracle"]
    A_gpu = solver._assemble_system().to_scipy()
    Ao = o.on_pattern(c["A"])
    fro = float(np.linalg.norm(A_gpu.data - Ao.data) / np.linalg.norm(Ao.data))
    x = solver.up.vector().get_local()
    sp_ = o.space
    errs = _field_errors(x, c["x"], [3 * sp_.N2] + [sp_.Nv] * o.nfields)
    solver.engine.close()
    trend = []
    for m in CPU_TREND_NS:
        t = run_cpu(config, m, rtol, 1, 1, maxit)
        trend.append({"n": m, "dofs": t["dofs"], "cpu_dof_s": t["dofs"] / t["sec"], "iterations": t["iters"]})
    trend.append({"n": n, "dofs": c["dofs"], "cpu_dof_s": c["dofs"] / c["sec"], "iterations": c["iters"]})
    matched = {"n": n, "dofs": c["dofs"], "steps": steps, "rtol": rtol, "gpu_dof_s": g["dof_per_s"],
               "gpu_ms_per_step": g["ms_per_step"], "gpu_iterations": g["krylov_iterations"],
               "cpu_dof_s": c["dofs"] / c["sec"], "cpu_s_per_step": c["sec"], "cpu_iterations": c["iters"],
               "cpu_cores": c["threads"], "ratio": g["dof_per_s"] / (c["dofs"] / c["sec"]),
               "A_frobenius_diff": fro, "field_L2_diff": errs,
               "amg_setup_s": {"gpu": g["amg_setup_s"], "cpu": c["setup_s"]}, "cpu_trend": trend}
    baseline = {"value": c["dofs"] / c["sec"], "unit": UNIT, "cores": c["threads"], "kind": "port",
                "sample": "%s family at n=%d (%d dofs), %d timed steps after %d warm-up, MINRES its %s, %.1f s per step"
                          % (config, n, c["dofs"], steps, warm, c["iters"], c["sec"])}
    return matched, baseline


# ------------------------------------------------------------------------------------------ main
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        c = run_cpu(args.config, args.cpu_n, args.rtol, max(1, args.steps), max(0, min(args.warmup, 1)), args.maxit)
        val = c["dofs"] / c["sec"]
        sample = "%s family at n=%d (%d dofs), %d step(s), MINRES its %s" % (args.config, args.cpu_n, c["dofs"],
                                                                             len(c["iters"]), c["iters"])
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": c["sec"] * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_string(args.config, args.n or CONFIG_N[args.config], args.rtol),
                           "sample": sample, "rtol": args.rtol,
                           "amg_setup_s_excluded": round(c["setup_s"], 2),
                           "note": "CPU restatement (OpenMP cell loop + OpenMP CSR products, oracle/) of the reference "
                                   "path; DOLFIN/PETSc are not installable offline.  Same workload family, tolerance "
                                   "and step definition as the b200 arm, which reports a matched run at this n "
                                   "(`matched`)"},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": c["threads"], "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    from waterscapes_b200.workloads import make_problem, CONFIGS
    from waterscapes_b200.mpet import MPETSolver, MPETTotalPressureSolver

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = CONFIGS[args.config]
    n = args.n or cfg["n"]
    partition = None
    grid = (1, 1, 1)
    if world > 1:
        # weak scaling on the REFINED CUBE of BASELINE.json configs[4]: the domain stays the cube, the mesh is
        # refined isotropically to n * world^(1/3) cells per edge (so every rank keeps ~n^3 cubes of 6 tets),
        # and is cut into a 3-D grid of bricks (+ one ghost cell layer around each brick)
        from waterscapes_b200.parallel import box_brick, box_slab, brick_grid, Partition
        ng = int(round(n * world ** (1.0 / 3.0)))
        L = 1.0 if args.config in ("cfg1", "cfg2") else 120.0
        grid = (1, 1, world) if args.partition == "slab" else brick_grid(world)
        mesh = box_brick((0.0, 0.0, 0.0), (L, L, L), ng, ng, ng, rank, world, grid=grid)
        problem, sp, init = make_problem(args.config, n, mesh=mesh)
        partition = Partition(rank, world)
    else:
        problem, sp, init = make_problem(args.config, n)
    sp = dict(sp, direct_solver=False, krylov_rtol=args.rtol, krylov_maxit=args.maxit)
    tp = args.formulation == "total-pressure"
    solver = (MPETTotalPressureSolver if tp else MPETSolver)(problem, sp, device=local_rank, partition=partition)
    eng = solver.engine
    S = eng.sizes
    if tp:      # initial network pressures = boundary data, total pressure 0 (the workload's own init targets MPETSolver)
        problem.time.assign(0.0)
        for i in range(int(problem.params["J"])):
            solver.up_.set_sub(solver._net_sub(i), problem.p_bar[i])
    else:
        init(solver)
    dt = sp["dt"]

    def one_step():
        solver.step(dt)
        solver.up_.assign(solver.up)

    # set-up outside the timed region on both arms: AMG hierarchy (the first solve builds it; timed on its own)
    barrier()
    t0 = time.perf_counter()
    one_step()
    barrier()
    first_step_s = time.perf_counter() - t0
    setup_s = float(solver.solver_monitor.get("pc_setup_s") or 0.0)
    for _ in range(max(args.warmup, 3) - 1):
        one_step()
    barrier()
    snap_state, snap_time = solver.up_.x.clone(), float(problem.time)

    # ---- device-resident timed region
    sampler = ClockSampler(local_rank)
    sampler.start()
    eng.profile(enable=0 if os.environ.get("MPET_BENCH_NOPROF") else 1)     # development aid: time the step without the profiler's events
    eng.launch_count(reset=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = []
    for _ in range(args.steps):
        one_step()
        iters.append(solver.solver_monitor["niter"][-1])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count()
    prof = eng.profile(enable=0)
    clocks = sampler.stop()
    rel_res, res_l2, rhs_l2 = true_residual(solver, partition, dist, world)

    # ---- end-to-end over the SAME time steps: host buffers in, host buffers out, through the public API
    N = S["N"]
    solver.up_.x.copy_(snap_state)
    problem.time.assign(snap_time)
    h_in = torch.empty(N, dtype=torch.float64).pin_memory()
    h_out = torch.empty(N, dtype=torch.float64).pin_memory()
    h_in.copy_(solver.up_.x)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    iters_e2e = []
    for _ in range(args.steps):
        solver.up_.x.copy_(h_in, non_blocking=True)          # H2D: previous state
        solver.step(dt)
        h_out.copy_(solver.up.x, non_blocking=True)           # D2H: new state
        torch.cuda.current_stream().synchronize()
        h_in.copy_(h_out)
        iters_e2e.append(solver.solver_monitor["niter"][-1])
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    t = torch.tensor([ms, ms_e2e, setup_s, first_step_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, setup_s, first_step_s = (float(v) for v in t)
    if world > 1:
        owned = torch.tensor([partition.n_owned_dofs()], dtype=torch.float64, device="cuda")
        dist.all_reduce(owned)
        total_dofs = int(owned.item())
    else:
        total_dofs = N
    ms_step = ms / args.steps
    value = total_dofs / (ms_step / 1e3)
    e2e_value = total_dofs / (ms_e2e / args.steps / 1e3)

    peak, peak_src = measured_peak()
    # algorithmic bytes of one node-blocked SpMV launch pair (DESIGN.md "SpMV"): every value once, one
    # column index per NODE pair, row pointers, x read once + y written once in the padded layout
    nint = 4 * S["N2"] + S["A"] * S["Nv"]
    spmv_bytes = (8 * S["nnz"] + 4 * (S["nnz22"] + 2 * S["nnz21"] + S["nnz11"]) + 8 * S["N"]
                  + 8 * (S["N2"] + S["Nv"]) + 16 * nint)
    csr_equiv_bytes = 12 * S["nnz"] + 20 * S["N"]
    spmv_ms = prof["spmv"]["ms"] / max(1, prof["spmv"]["count"])
    pc_ms = prof["pc"]["ms"] / max(1, prof["pc"]["count"])
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else 0.0
    traffic = ncu_traffic() or {}
    pc_bytes = eng.pc_bytes()
    total_its = max(1, sum(iters))
    roofline = {"bound": "hbm", "kernel": "k_block_rows_pipe<3,A,1,16> + <A,A,0,32> (node-blocked SpMV of the block system inside MINRES, staged.cu)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "csr_equivalent_gbs": csr_equiv_bytes / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else 0.0,
                "note": "bytes are the node-blocked format's own (one column index per 9 values of a node pair); "
                        "csr_equivalent_gbs counts SURVEY 8(d)'s scalar-CSR bytes for the same product",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": spmv_bytes,
                "avg_launch_ms": spmv_ms, "launches_timed": prof["spmv"]["count"],
                "traffic": traffic.get("dram_bytes_per_launch") if world == 1 and n == traffic.get("n") else None,
                "traffic_source": traffic.get("source") if world == 1 and n == traffic.get("n") else None,
                "preconditioner": {"kernel": "one block-diagonal AMG application (k_spmm_pipe / k_cheb_first / k_dense_apply passes over all levels, amg.cu)",
                                   "algorithmic_bytes_per_application": pc_bytes, "avg_application_ms": pc_ms,
                                   "achieved": pc_bytes / (pc_ms * 1e-3) / 1e9 if pc_ms > 0 else 0.0,
                                   "frac": pc_bytes / (pc_ms * 1e-3) / 1e9 / peak if pc_ms > 0 else 0.0,
                                   "applications_timed": prof["pc"]["count"]},
                "share_of_step": {"spmv": prof["spmv"]["ms"] / ms, "preconditioner": prof["pc"]["ms"] / ms,
                                  "assemble_lhs": prof["assemble"]["ms"] / ms, "rhs_prev": prof["rhs"]["ms"] / ms,
                                  "halo_and_collectives_main_stream": prof["comm"]["ms"] / ms,
                                  # only with MPET_PC_STREAMS=0 (serial preconditioner blocks): displacement / P1 fields
                                  "pc_displacement_serial": prof["pc_u"]["ms"] / ms, "pc_p1_fields_serial": prof["pc_p"]["ms"] / ms},
                "ms_per_iteration": ms / total_its, "launches_per_iteration": launches / total_its,
                "comm_ops_timed": prof["comm"]["count"]}

    parity = None
    try:
        parity = parity_distributed(args.config, world, rank, local_rank, dist) if world > 1 else \
            parity_single(args.config, local_rank)
    except Exception as e:                                    # a failed check is reported, never hidden
        parity = {"error": "%s: %s" % (type(e).__name__, e)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"formulation": args.formulation,
                       "workload": workload_string(args.config, n, args.rtol),
                       "dofs_per_gpu": N, "cells_per_gpu": S["Nc"], "nnz_per_gpu": S["nnz"], "rtol": args.rtol,
                       "krylov_iterations": iters,
                       "krylov_iterations_e2e": iters_e2e,
                       "parallelism": "1 GPU" if world == 1 else "cube refined to %d^3 cubes, %dx%dx%d bricks (own cells + 1 ghost layer), %s halo exchange + all-reduced dots, distributed V-cycles" % ((int(round(n * world ** (1.0 / 3.0))),) + tuple(grid) + (eng.comm_kind(),)),
                       "total_dofs": total_dofs,
                       "l2": "inputs (%.1f GB matrix) larger than the 126 MB L2" % (12 * S["nnz"] / 1e9),
                       "amg_setup_s_excluded": round(setup_s, 2), "first_step_s": round(first_step_s, 2)},
            "true_residual": {"rel": rel_res, "residual_l2": res_l2, "rhs_l2": rhs_l2,
                              "definition": "||b - A x||_2 / ||b||_2 of the last timed step, free owned rows, b after the symmetric Dirichlet elimination"},
            "parity": parity,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": 8 * N, "d2h_bytes_per_step": 8 * N},
            "gpu_launches": launches,
            "roofline": roofline}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        eng_main = eng
        matched, baseline = matched_block(args.config, args.cpu_n, args.rtol, args.maxit, local_rank)
        line["matched"] = matched
        line["cpu_baseline"] = baseline
    if rank == 0 and world == 1 and not args.no_configs:
        # the other BASELINE.json configurations at their stated sizes (parity-test cases; timed here for the record)
        table = []
        for name in ("cfg1", "cfg2", "cfg3", "cfg4"):
            if name == args.config:
                continue
            try:
                table.append(gpu_time_config(name, CONFIGS[name]["n"], args.rtol, args.maxit, 3, 2, local_rank))
            except Exception as e:
                table.append({"config": name, "error": "%s: %s" % (type(e).__name__, e)})
        line["configs"] = table
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
