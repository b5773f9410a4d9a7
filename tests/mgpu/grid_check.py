"""Nodes x slabs on GPUs (run under torchrun on 4 GPUs of one node):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 \
        tests/mgpu/grid_check.py [n]

2-D process grid from ``parallel.cartesian_comms(2, 2)``: the node-parallel sweeper (one collocation node per outer rank,
f all-gathered over NCCL) on slab-decomposed fields (persistent peer-memory CG inside each node's pair of GPUs).  Rank 0
also runs the same step on one GPU; checks: identical SDC iteration counts, end value <= 1e-10 relative."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 63
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    from pysdc_b200.controller import controller_nonMPI
    from pysdc_b200.datatypes import mesh
    from pysdc_b200.parallel import cartesian_comms
    from pysdc_b200.problems import heatNd_unforced
    from pysdc_b200.stats import get_sorted
    from pysdc_b200.sweepers import generic_implicit, generic_implicit_MPI

    n_nodes = 2
    node_comm, space_comm = cartesian_comms(n_nodes, world // n_nodes)
    pp = dict(nvars=(n, n, n), nu=0.1, freq=(1, 1, 1), bc="dirichlet-zero", solver_type="CG", lintol=1e-12, liniter=10000)
    sp = dict(num_nodes=n_nodes, quad_type="RADAU-RIGHT", QI="MIN-SR-NS", initial_guess="spread")
    u_g = np.random.default_rng(1234).standard_normal((n, n, n))

    def run(problem_params, sweeper_class, sweeper_params):
        c = controller_nonMPI(1, {"logger_level": 40}, dict(
            problem_class=heatNd_unforced, problem_params=problem_params, sweeper_class=sweeper_class,
            sweeper_params=sweeper_params, level_params=dict(dt=1e-3, restol=1e-9), step_params=dict(maxiter=30)))
        P = c.MS[0].levels[0].prob
        u0 = P.dtype_u(P.init)
        u0[:] = u_g
        uend, stats = c.run(u0=u0, t0=0.0, Tend=2e-3)
        return uend, [int(v) for _, v in get_sorted(stats, type="niter", sortby="time")]

    uend, niter = run(dict(pp, comm=space_comm), generic_implicit_MPI, dict(sp, comm=node_comm))
    got = uend.gather()
    ok, info = True, {}
    if rank == 0:
        mesh.comm = None
        ref, niter_ref = run(pp, generic_implicit, dict(sp))
        ref = ref.get()
        info = dict(niter=(niter, niter_ref), uend_vs_1gpu=float(np.max(np.abs(got - ref)) / np.max(np.abs(ref))))
        ok = niter == niter_ref and info["uend_vs_1gpu"] < 1e-10
    flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"grid_check n={n} world={world} ({n_nodes} nodes x {world // n_nodes} slabs): {'OK' if ok else 'MISMATCH'} {info}", flush=True)
    dist.destroy_process_group()
    return 0 if float(flag.item()) == 1.0 else 1


if __name__ == "__main__":
    sys.exit(main())
