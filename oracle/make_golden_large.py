#!/usr/bin/env python
"""Golden fixtures at the sizes BASELINE.md section 4 names, from the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE.  Companion of ``make_golden.py`` for the slow cases: config 2 (``heatNd_forced`` 2-D, IMEX LU) at
511^2, 1023^2 and the full 2047^2; config 4 (``allencahn_fullyimplicit``) at 256^2 and 512^2, where the inner CG runs
into ``lin_maxiter``.  A fixture keeps the iteration counts, the work counters, the residual after every iteration,
the max-norm of ``uend`` and a subsample of ``uend`` (every ``subsample``-th point) instead of the full field.

    python oracle/make_golden_large.py config2_511 config2_1023 config4_256 config4_512     # minutes
    python oracle/make_golden_large.py config2_2047                                          # ~40 min of CPU
"""
import json
import sys
import time

import numpy as np

import make_golden as mg


def run_large(name, spec, subsample):
    d = mg.make_description(spec)
    t_start = time.perf_counter()
    c = mg.controller_nonMPI(num_procs=1, controller_params={"logger_level": 40, "hook_class": [mg.LogWork]},
                             description=d)
    P = c.MS[0].levels[0].prob
    u0 = mg.initial_value(P, spec)
    uend, stats = c.run(u0=u0, t0=spec["t0"], Tend=spec["Tend"])
    wall = time.perf_counter() - t_start
    gs = mg.get_sorted
    niter = [int(v) for _, v in gs(stats, type="niter", sortby="time")]
    times = [float(t) for t, _ in gs(stats, type="niter", sortby="time")]
    res = np.full((len(times), max(niter)), np.nan)
    for i, t in enumerate(times):
        r = [float(v) for _, v in gs(stats, time=t, type="residual_post_iteration", sortby="iter")]
        res[i, : len(r)] = r
    arrays = dict(niter=np.array(niter), times=np.array(times), residuals=res, uend_maxabs=np.array(float(abs(uend))),
                  uend_sub=np.asarray(uend)[::subsample, ::subsample].copy(), wall_seconds=np.array(wall))
    for key in P.work_counters:
        arrays["work_" + key] = np.array([int(v) for _, v in gs(stats, type="work_" + key, sortby="time")])
    if spec["problem"] != "allencahn_fullyimplicit":
        arrays["err_vs_exact"] = np.array(float(abs(P.u_exact(spec["Tend"]) - uend)))
    mg.save(name, dict(spec, subsample=subsample), **arrays)
    print(f"  {name}: niter {niter}, work { {k: arrays['work_' + k].tolist() for k in P.work_counters} }, "
          f"|uend| {float(abs(uend))!r}, wall {wall:.0f} s", flush=True)


def config2(n, steps=2):
    spec = dict(problem="heatNd_forced", sweeper="imex_1st_order",
                problem_params=dict(nvars=[n, n], nu=0.1, freq=[4, 4], bc="dirichlet-zero", solver_type="CG",
                                    lintol=1e-12, liniter=10000),
                sweeper_params=dict(num_nodes=4, quad_type="RADAU-RIGHT", QI="LU"),
                level_params=dict(dt=0.1, restol=1e-10), step_params=dict(maxiter=50),
                t0=0.0, Tend=0.1 * steps, u0="exact")
    run_large(f"run_config2_heat2d_imex_lu_{n}", spec, subsample=max(1, (n + 1) // 64))


def config4(n, steps=2):
    spec = dict(problem="allencahn_fullyimplicit", sweeper="generic_implicit",
                problem_params=dict(nvars=[n, n], nu=2, eps=0.04, newton_maxiter=100, newton_tol=1e-9,
                                    lin_tol=1e-10, lin_maxiter=100, radius=0.25),
                sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU", initial_guess="zero"),
                level_params=dict(dt=1e-3, restol=1e-8), step_params=dict(maxiter=50),
                t0=0.0, Tend=1e-3 * steps, u0="exact")
    run_large(f"run_config4_allencahn_gi_lu_{n}", spec, subsample=max(1, n // 64))


def config3(n, K=4):
    """Config 3 scaled down with the bench settings (restol=-1, K sweeps, seeded random field)."""
    spec = dict(problem="heatNd_unforced", sweeper="generic_implicit",
                problem_params=dict(nvars=[n, n, n], nu=0.1, freq=[1, 1, 1], bc="dirichlet-zero", solver_type="CG",
                                    lintol=1e-12, liniter=10000),
                sweeper_params=dict(num_nodes=4, quad_type="RADAU-RIGHT", QI="MIN-SR-NS", initial_guess="spread"),
                level_params=dict(dt=1e-3, restol=-1), step_params=dict(maxiter=K),
                t0=0.0, Tend=1e-3, u0="random", seed=1234)
    d = mg.make_description(spec)
    c = mg.controller_nonMPI(num_procs=1, controller_params={"logger_level": 40, "hook_class": [mg.LogWork]}, description=d)
    P = c.MS[0].levels[0].prob
    u0 = mg.initial_value(P, spec)
    uend, stats = c.run(u0=u0, t0=0.0, Tend=1e-3)
    gs = mg.get_sorted
    niter = [int(v) for _, v in gs(stats, type="niter", sortby="time")]
    res = np.array([[float(v) for _, v in gs(stats, type="residual_post_iteration", sortby="iter")]])
    sub = max(1, (n + 1) // 32)
    mg.save(f"run_config3_heat3d_gi_minsrns_{n}_K{K}", dict(spec, subsample=sub), niter=np.array(niter), times=np.array([0.0]),
            residuals=res, uend_maxabs=np.array(float(abs(uend))), uend_sub=np.asarray(uend)[::sub, ::sub, ::sub].copy(),
            work_CG=np.array([int(v) for _, v in gs(stats, type="work_CG", sortby="time")]))
    print(f"  config3 {n}: niter {niter}, |uend| {float(abs(uend))!r}", flush=True)


FAMILIES = {"config3_127": lambda: config3(127), "config2_511": lambda: config2(511), "config2_1023": lambda: config2(1023),
            "config2_2047": lambda: config2(2047), "config4_256": lambda: config4(256),
            "config4_512": lambda: config4(512), "config4_1024": lambda: config4(1024)}

if __name__ == "__main__":
    for w in sys.argv[1:]:
        FAMILIES[w]()
