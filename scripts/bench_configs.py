#!/usr/bin/env python
"""Secondary measurements: BASELINE.json configs 2 and 4 at full size on one GPU (not the headline bench line).

    python scripts/bench_configs.py [--steps 2]

Prints one JSON line per config: SDC iterations per step, solver work counters, wall per step, DOF-node updates/s
(N * M * total sweeps / time, the BASELINE metric)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pysdc_b200.controller import LogWork, controller_nonMPI  # noqa: E402
from pysdc_b200.problems import allencahn_fullyimplicit, heatNd_forced  # noqa: E402
from pysdc_b200.stats import get_sorted  # noqa: E402
from pysdc_b200.sweepers import generic_implicit, imex_1st_order  # noqa: E402

CONFIGS = {
    "config2_heat2d_forced_2047_imex_LU": dict(
        problem_class=heatNd_forced,
        problem_params=dict(nvars=(2047, 2047), nu=0.1, freq=(4, 4), bc="dirichlet-zero", solver_type="CG", lintol=1e-12,
                            liniter=10000),
        sweeper_class=imex_1st_order, sweeper_params=dict(num_nodes=4, quad_type="RADAU-RIGHT", QI="LU"),
        level_params=dict(dt=0.1, restol=1e-10), step_params=dict(maxiter=50)),
    "config4_allencahn_2048_newton_LU": dict(
        problem_class=allencahn_fullyimplicit,
        problem_params=dict(nvars=(2048, 2048), nu=2, eps=0.04, newton_maxiter=100, newton_tol=1e-9, lin_tol=1e-10,
                            lin_maxiter=100, radius=0.25),
        sweeper_class=generic_implicit,
        sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU", initial_guess="zero"),
        level_params=dict(dt=1e-3, restol=1e-8), step_params=dict(maxiter=50)),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    for name, d in CONFIGS.items():
        c = controller_nonMPI(1, {"logger_level": 40, "hook_class": [LogWork]}, d)
        L = c.MS[0].levels[0]
        P = L.prob
        u0 = P.u_exact(0.0)
        dt = d["level_params"]["dt"]
        c.run(u0=u0, t0=0.0, Tend=dt)  # warm-up (workspace allocation, module load)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        uend, stats = c.run(u0=u0, t0=0.0, Tend=args.steps * dt)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        niter = [int(v) for _, v in get_sorted(stats, type="niter", sortby="time")]
        work = {k[len("work_"):]: [int(v) for _, v in get_sorted(stats, type=k, sortby="time")]
                for k in sorted({e.type for e in stats if str(e.type).startswith("work_")})}
        times = [t for t, _ in get_sorted(stats, type="niter", sortby="time")]
        hist = [[float(v) for _, v in get_sorted(stats, time=t, type="residual_post_iteration", sortby="iter")][-6:]
                for t in times]
        N, M = int(np.prod(P.nvars)), L.sweep.coll.num_nodes
        print(json.dumps(dict(config=name, steps=args.steps, niter=niter, work=work, wall_s=wall,
                              ms_per_step=1e3 * wall / args.steps, dof_node_updates_per_s=N * M * sum(niter) / wall,
                              uend_maxnorm=float(abs(uend)), last_residuals=hist)), flush=True)


if __name__ == "__main__":
    main()
