"""Multi-GPU parity check of the slab-decomposed path (run under torchrun on >= 2 GPUs of one node):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/mgpu/slab_check.py [n]

Every rank solves its slab of a seeded 3-D heat problem with the persistent peer-memory CG and runs a full SDC step;
rank 0 also runs the same problem on one GPU and with the CPU oracle.  Checks: eval_f bitwise equal to the single-GPU
kernel, solve and end-of-step solution <= 1e-10 relative to both, identical SDC iteration counts."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 63
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from pysdc_b200.controller import controller_nonMPI
    from pysdc_b200.datatypes import mesh
    from pysdc_b200.parallel import SlabComm
    from pysdc_b200.problems import heatNd_unforced
    from pysdc_b200.stats import get_sorted
    from pysdc_b200.sweepers import generic_implicit

    comm = SlabComm()
    pp = dict(nvars=(n, n, n), nu=0.1, freq=(1, 1, 1), bc="dirichlet-zero", solver_type="CG", lintol=1e-12, liniter=10000)
    sweeper_params = dict(num_nodes=4, quad_type="RADAU-RIGHT", QI="MIN-SR-NS", initial_guess="spread")
    rng = np.random.default_rng(1234)
    u_g, rhs_g = rng.standard_normal((n, n, n)), rng.standard_normal((n, n, n))
    factors = [2.5e-4 * (m + 1) for m in range(4)]

    def to_mesh(P, arr):
        m = P.dtype_u(P.init)
        m[:] = arr
        return m

    # ---- slab run ---------------------------------------------------------------------------------------------------
    P = heatNd_unforced(**pp, comm=comm)
    f_slab = P.eval_f(to_mesh(P, u_g), 0.0).gather()
    xs = [to_mesh(P, u_g) for _ in factors]
    P.solve_system_batch([to_mesh(P, rhs_g) for _ in factors], factors, xs)
    sol_slab = [x.gather() for x in xs]
    cg_slab = P.work_counters["CG"].niter
    # the same solves with the polynomial preconditioner (z halo planes travel through peer memory as well)
    Pp = heatNd_unforced(**pp, comm=comm, preconditioner="chebyshev")
    xp = [to_mesh(Pp, u_g) for _ in factors]
    Pp.solve_system_batch([to_mesh(Pp, rhs_g) for _ in factors], factors, xp)
    sol_pc = [x.gather() for x in xp]
    cg_pc = Pp.work_counters["CG"].niter

    def run(pp_run):
        c = controller_nonMPI(1, {"logger_level": 40}, dict(
            problem_class=heatNd_unforced, problem_params=pp_run, sweeper_class=generic_implicit,
            sweeper_params=dict(sweeper_params), level_params=dict(dt=1e-3, restol=1e-8), step_params=dict(maxiter=50)))
        Pr = c.MS[0].levels[0].prob
        uend, stats = c.run(u0=to_mesh(Pr, u_g), t0=0.0, Tend=2e-3)
        return uend, [v for _, v in get_sorted(stats, type="niter")]

    uend_slab, niter_slab = run(dict(pp, comm=comm))
    uend_slab = uend_slab.gather()

    # forced heat, IMEX sweeper with LU (sequential node solves, one system per launch) on slabs
    from pysdc_b200.problems import heatNd_forced
    from pysdc_b200.sweepers import imex_1st_order

    def run_imex(extra):
        c = controller_nonMPI(1, {"logger_level": 40}, dict(
            problem_class=heatNd_forced, problem_params=dict(pp, freq=(2, 2, 2), **extra), sweeper_class=imex_1st_order,
            sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU"),
            level_params=dict(dt=5e-3, restol=1e-9), step_params=dict(maxiter=50)))
        Pr = c.MS[0].levels[0].prob
        uend, stats = c.run(u0=Pr.u_exact(0.0), t0=0.0, Tend=1e-2)
        return uend, [v for _, v in get_sorted(stats, type="niter")]

    uend_imex_slab, niter_imex_slab = run_imex(dict(comm=comm))
    uend_imex_slab = uend_imex_slab.gather()
    torch.cuda.synchronize()
    dist.barrier()

    # ---- single-GPU and oracle runs on rank 0 ----------------------------------------------------------------------
    ok = True
    if rank == 0:
        mesh.comm = None  # class-level communicator (mesh.py:46): back to serial fields
        P1 = heatNd_unforced(**pp)
        f_one = P1.eval_f(to_mesh(P1, u_g), 0.0).get()
        x1 = [to_mesh(P1, u_g) for _ in factors]
        P1.solve_system_batch([to_mesh(P1, rhs_g) for _ in factors], factors, x1)
        uend_one, niter_one = run(dict(pp))
        uend_one = uend_one.get()
        uend_imex_one, niter_imex_one = run_imex({})
        uend_imex_one = uend_imex_one.get()
        rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))  # noqa: E731
        checks = {"eval_f bitwise": bool(np.array_equal(f_slab, f_one)),
                  "solve vs 1 GPU": max(rel(a, b.get()) for a, b in zip(sol_slab, x1)),
                  "CG its slab / 1 GPU": (cg_slab, P1.work_counters["CG"].niter),
                  "preconditioned solve vs 1 GPU": max(rel(a, b.get()) for a, b in zip(sol_pc, x1)),
                  "preconditioned its": cg_pc,
                  "uend vs 1 GPU": rel(uend_slab, uend_one), "niter": (niter_slab, niter_one),
                  "imex uend vs 1 GPU": rel(uend_imex_slab, uend_imex_one), "imex niter": (niter_imex_slab, niter_imex_one)}
        ok = checks["eval_f bitwise"] and checks["solve vs 1 GPU"] < 1e-10 and checks["uend vs 1 GPU"] < 1e-10 \
            and niter_slab == niter_one and abs(cg_slab - P1.work_counters["CG"].niter) <= 4 \
            and checks["preconditioned solve vs 1 GPU"] < 1e-10 and cg_pc < 0.8 * cg_slab \
            and checks["imex uend vs 1 GPU"] < 1e-10 and niter_imex_slab == niter_imex_one
        if n <= 63:
            import sdc_oracle

            spec = dict(problem="heatNd_unforced", sweeper="generic_implicit",
                        problem_params=dict(pp, nvars=[n] * 3, freq=[1, 1, 1]), sweeper_params=sweeper_params,
                        level_params=dict(dt=1e-3, restol=1e-8), step_params=dict(maxiter=50), t0=0.0, Tend=2e-3,
                        u0="random", seed=1234)
            ref = sdc_oracle.run_sdc(spec, u0=u_g.copy())
            checks["uend vs oracle"] = rel(uend_slab, ref["uend"])
            checks["niter oracle"] = ref["niter"]
            ok = ok and checks["uend vs oracle"] < 1e-10 and ref["niter"] == niter_slab
        print(f"slab_check n={n} world={world}: {'OK' if ok else 'FAILED'} {checks}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
