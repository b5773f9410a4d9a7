#!/usr/bin/env python
"""How firmly does the UNMODIFIED reference itself hold its SDC iteration counts at the BASELINE sizes?

TEST INFRASTRUCTURE (build container only; imports /root/reference).  At 2047^2 (config 2) and 1023^2/511^2 PFASST
(config 5) the runs stagnate at the accuracy the inner CG (lintol = 1e-12 on the 2-norm) can deliver, a few 1e-10 in the
max-norm residual, right at restol = 1e-10: the iteration at which ``residual <= restol`` first holds
(convergence_controller_classes/check_convergence.py:75-76) is then decided by rounding.  This script documents that
with the reference's own classes: it re-runs them with inputs perturbed at rounding level —

  * ``lintol * (1 +- 1e-3)`` / ``(1 +- 1e-2)``: moves the CG stopping iteration of a few node solves by one,
  * ``u0 * (1 + k ulp)``: a relative perturbation of the initial value by a few units in the last place,

and records the iteration counts and the residual histories.  The result is committed as
``tests/golden/sensitivity_*.json``; the parity tests accept a differing count only at (step, iteration) pairs that this
record shows the reference flipping on.

    python oracle/sensitivity.py config2 lintol 0.999          # one perturbed run -> /tmp/gold/sens_config2_lintol_0.999.json
    python oracle/sensitivity.py collect                       # merge /tmp/gold/sens_*.json into tests/golden/
"""
import glob
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TMP = os.environ.get("SENS_TMP", "/tmp/gold")


def _perturb_u0(u0, seed):
    """u0 * (1 + k * 2^-52), k uniform in {-2..2}: a perturbation of a few units in the last place."""
    rng = np.random.default_rng(seed)
    u0[:] = np.asarray(u0) * (1.0 + rng.integers(-2, 3, size=u0.shape) * 2.0**-52)
    return u0


def _histories(stats, get_sorted):
    niter = [int(v) for _, v in get_sorted(stats, type="niter", sortby="time")]
    hist = {}
    for k, v in stats.items():
        if k.type == "residual_post_iteration":
            hist.setdefault(round(k.time, 10), {})[k.iter] = float(v)
    res = [[hist[t].get(i + 1) for i in range(max(hist[t]))] for t in sorted(hist)]
    return niter, res


def config2(kind, value, n=2047, steps=1):
    import make_golden as mg

    lintol = 1e-12 * (float(value) if kind == "lintol" else 1.0)
    spec = dict(problem="heatNd_forced", sweeper="imex_1st_order",
                problem_params=dict(nvars=[n, n], nu=0.1, freq=[4, 4], bc="dirichlet-zero", solver_type="CG",
                                    lintol=lintol, liniter=10000),
                sweeper_params=dict(num_nodes=4, quad_type="RADAU-RIGHT", QI="LU"),
                level_params=dict(dt=0.1, restol=1e-10), step_params=dict(maxiter=50),
                t0=0.0, Tend=0.1 * steps, u0="exact")
    c = mg.controller_nonMPI(num_procs=1, controller_params={"logger_level": 40, "hook_class": [mg.LogWork]},
                             description=mg.make_description(spec))
    P = c.MS[0].levels[0].prob
    u0 = P.u_exact(0.0)
    if kind == "u0ulp":
        u0 = _perturb_u0(u0, int(value))
    t0 = time.perf_counter()
    uend, stats = c.run(u0=u0, t0=0.0, Tend=spec["Tend"])
    niter, res = _histories(stats, mg.get_sorted)
    cg = [int(v) for _, v in mg.get_sorted(stats, type="work_CG", sortby="time")]
    return dict(config="config2", n=n, steps=steps, perturbation={kind: value}, niter=niter, work_CG=cg, residuals=res,
                uend_maxabs=float(abs(uend)), wall_seconds=time.perf_counter() - t0)


def config5(kind, value, n=1023, nprocs=8):
    import make_golden as mg

    lintol = 1e-12 * (float(value) if kind == "lintol" else 1.0)
    pp = dict(nvars=[(n, n), (n // 2, n // 2)], nu=0.1, freq=(4, 4), bc="dirichlet-zero", solver_type="CG",
              lintol=lintol, liniter=10000)
    d = dict(problem_class=mg.heatNd_forced, problem_params=pp, sweeper_class=mg.imex_1st_order,
             sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU"),
             level_params=dict(dt=0.25, restol=1e-10), step_params=dict(maxiter=50),
             space_transfer_class=mg.mesh_to_mesh, space_transfer_params=dict(rorder=2, iorder=6))
    c = mg.controller_nonMPI(num_procs=nprocs, controller_params=dict(logger_level=40, predict_type="pfasst_burnin"),
                             description=d)
    P = c.MS[0].levels[0].prob
    u0 = P.u_exact(0.0)
    if kind == "u0ulp":
        u0 = _perturb_u0(u0, int(value))
    t0 = time.perf_counter()
    uend, stats = c.run(u0=u0, t0=0.0, Tend=0.25 * nprocs)
    niter, res = _histories(stats, mg.get_sorted)
    return dict(config="config5", n=n, num_procs=nprocs, perturbation={kind: value}, niter=niter, residuals=res,
                uend_maxabs=float(abs(uend)), wall_seconds=time.perf_counter() - t0)


def collect():
    out = {}
    for path in sorted(glob.glob(os.path.join(TMP, "sens_*.json"))):
        with open(path) as f:
            r = json.load(f)
        out.setdefault(r["config"], []).append(r)
    golden = {"config2": "run_config2_heat2d_imex_lu_2047", "config5": "pfasst_config5_1023_p8"}
    for cfg, runs in out.items():
        gdir = os.path.join(os.path.dirname(HERE), "tests", "golden")
        base = np.load(os.path.join(gdir, golden[cfg] + ".npz"))["niter"].tolist()  # the unperturbed reference run
        for r in runs:
            r["niter_unperturbed"] = base
            r["residuals"] = [h[-4:] for h in r["residuals"]]  # the last iterations: where the stopping test is decided
        dst = os.path.join(gdir, f"sensitivity_{cfg}.json")
        with open(dst, "w") as f:
            json.dump(dict(note="unmodified reference with rounding-level input perturbations (oracle/sensitivity.py); "
                                "niter_unperturbed = the reference's own answer on the unperturbed inputs",
                           runs=runs), f, indent=1)
        print("wrote", dst, [(r["perturbation"], r["niter"]) for r in runs])


if __name__ == "__main__":
    if sys.argv[1] == "collect":
        collect()
    else:
        cfg, kind, value = sys.argv[1:4]
        extra = [int(a) for a in sys.argv[4:]]
        r = {"config2": config2, "config5": config5}[cfg](kind, value, *extra)
        os.makedirs(TMP, exist_ok=True)
        tag = "_".join([cfg, kind, str(value)] + [str(e) for e in extra])
        with open(os.path.join(TMP, f"sens_{tag}.json"), "w") as f:
            json.dump(r, f)
        print(tag, r["niter"], f"{r['wall_seconds']:.0f} s", flush=True)
