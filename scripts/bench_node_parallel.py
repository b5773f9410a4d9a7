#!/usr/bin/env python
"""BASELINE config 3 (3-D heat n^3, M = 4 MIN-SR-NS, K = 4 sweeps) "parallel across the method": one collocation node per
GPU with ``sweepers.generic_implicit_MPI`` (the B200 form of the reference's generic_implicit_MPI sweeper), launched as

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port P \
        scripts/bench_node_parallel.py [--n 511] [--steps 5] [--warmup 3]

Same timing rules as bench.py (barrier + synchronize on both sides, CUDA events, max over ranks) and the same answer
check against the committed single-GPU record of the workload.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    rank, local_rank, world = (int(os.environ[k]) for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from pysdc_b200 import problems, sweepers
    from pysdc_b200.controller import controller_nonMPI
    from pysdc_b200.parallel import TorchComm
    from pysdc_b200.stats import get_sorted

    w = bench.workload(3, args.n)
    assert world == w["M"], f"one rank per collocation node: launch {w['M']} ranks"
    d = bench.description(w, dict(heatNd_unforced=problems.heatNd_unforced, generic_implicit=sweepers.generic_implicit_MPI))
    d["sweeper_params"]["comm"] = TorchComm()
    ctrl = controller_nonMPI(num_procs=1, controller_params={"logger_level": 40}, description=d)
    L = ctrl.MS[0].levels[0]
    P = L.prob
    n = w["n"]
    u0 = P.dtype_u(P.init)
    rng = np.random.default_rng(w["seed"])  # every rank holds the whole field (nodes are distributed, space is not)
    host = torch.empty((n, n, n), dtype=torch.float64).pin_memory()
    for z in range(n):
        host.numpy()[z] = rng.standard_normal((n, n))
    u0.data.copy_(host)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        ctrl.run(u0=u0, t0=0.0, Tend=w["dt"])
    P.solve_log = []
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        uend, stats = ctrl.run(u0=u0, t0=0.0, Tend=w["dt"])
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    solve_ms = [a.elapsed_time(b) for a, b, _ in P.solve_log]
    its = [int(c.sum().item()) for _, _, c in P.solve_log]
    times = [tt for tt, _ in get_sorted(stats, type="niter", sortby="time")]
    res = [[float(v) for _, v in get_sorted(stats, time=tt, type="residual_post_iteration", sortby="iter")] for tt in times]
    niter = [int(v) for _, v in get_sorted(stats, type="niter", sortby="time")]
    check = bench.check_answer(w, 1, res, niter, float(abs(uend)))
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(rank=rank, solve_ms_per_launch=float(np.mean(solve_ms)),
                                          cg_it_per_solve=float(np.mean(its)), check=check["status"]))
    if rank == 0:
        K = w["step_params"]["maxiter"]
        print(json.dumps(dict(metric=w["metric"], value=w["ndof"] * w["M"] * K / (ms * 1e-3), unit="DOF-node updates/s",
                              n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms, higher_is_better=True,
                              scaling="strong", dtype="f64", data="synthetic",
                              config=dict(workload=w["name"], parallelism=f"nodes{world}: one collocation node per GPU, "
                                          "f all-gathered once per sweep over NCCL", inputs=w["inputs"]),
                              check=check, per_rank=gathered)), flush=True)
    dist.destroy_process_group()
    return 0 if check["status"] == "ok" else 1


if __name__ == "__main__":
    sys.exit(main())
