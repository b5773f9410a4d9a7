#!/usr/bin/env python
"""Generate the golden fixtures under ``tests/golden/`` from the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Runs only in the build container, where ``/root/reference`` (pySDC v5.6) is mounted;
it cannot run on the GPU box.  The reference is imported as-is with ``oracle/qmat_shim`` standing in for the
absent third-party ``qmat`` package.  Every fixture stores the inputs needed to replay the case (parameters as a
JSON string, seeds) and the reference outputs (iteration counts, residual histories, work counters, fields).

    python oracle/make_golden.py            # regenerate everything
    python oracle/make_golden.py sweep_gi   # one family

The known answers the reference itself asserts are re-asserted here before anything is written:
``tutorial/step_3/A_getting_statistics.py:43`` (12 iterations on every step) and
``tutorial/step_8/A_visualize_residuals.py:56-58`` (PFASST: 7 iterations on every step).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PYSDC_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "qmat_shim"))
sys.path.insert(0, REF)
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

from pySDC.core.step import Step  # noqa: E402
from pySDC.helpers.stats_helper import get_sorted  # noqa: E402
from pySDC.implementations.controller_classes.controller_nonMPI import controller_nonMPI  # noqa: E402
from pySDC.implementations.hooks.log_work import LogWork  # noqa: E402
from pySDC.implementations.problem_classes.AllenCahn_2D_FD import (  # noqa: E402
    allencahn_fullyimplicit, allencahn_multiimplicit, allencahn_multiimplicit_v2, allencahn_semiimplicit,
    allencahn_semiimplicit_v2)
from pySDC.implementations.problem_classes.AdvectionEquation_ND_FD import advectionNd  # noqa: E402
from pySDC.implementations.problem_classes.HeatEquation_ND_FD import heatNd_forced, heatNd_unforced  # noqa: E402
from pySDC.implementations.sweeper_classes.generic_implicit import generic_implicit  # noqa: E402
from pySDC.implementations.sweeper_classes.imex_1st_order import imex_1st_order  # noqa: E402
from pySDC.implementations.sweeper_classes.multi_implicit import multi_implicit  # noqa: E402
from pySDC.implementations.transfer_classes.TransferMesh import mesh_to_mesh  # noqa: E402

PROBLEMS = {"heatNd_unforced": heatNd_unforced, "heatNd_forced": heatNd_forced, "advectionNd": advectionNd,
            "allencahn_fullyimplicit": allencahn_fullyimplicit, "allencahn_semiimplicit": allencahn_semiimplicit,
            "allencahn_semiimplicit_v2": allencahn_semiimplicit_v2, "allencahn_multiimplicit": allencahn_multiimplicit,
            "allencahn_multiimplicit_v2": allencahn_multiimplicit_v2}
SWEEPERS = {"generic_implicit": generic_implicit, "imex_1st_order": imex_1st_order, "multi_implicit": multi_implicit}


def _jsonable(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, tuple):
            v = list(v)
        out[k] = v
    return out


def save(name, spec, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, spec=json.dumps(spec), **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KiB)")


def initial_value(P, spec):
    """u0 as specified: 'exact' -> u_exact(t0); 'random' -> default_rng(seed).standard_normal(nvars)."""
    if spec["u0"] == "exact":
        return P.u_exact(spec["t0"])
    u0 = P.u_init
    u0[:] = np.random.default_rng(spec["seed"]).standard_normal(u0.shape)
    return u0


def make_description(spec):
    pp = dict(spec["problem_params"])
    for k in ("nvars", "freq"):
        if isinstance(pp.get(k), list):
            pp[k] = tuple(pp[k])
    d = {
        "problem_class": PROBLEMS[spec["problem"]],
        "problem_params": pp,
        "sweeper_class": SWEEPERS[spec["sweeper"]],
        "sweeper_params": dict(spec["sweeper_params"]),
        "level_params": dict(spec["level_params"]),
        "step_params": dict(spec["step_params"]),
    }
    return d


def gmres_fixtures():
    """solver_type='GMRES' (generic_ND_FD.py:241-250) and advectionNd (AdvectionEquation_ND_FD.py): operator vectors for
    every stencil family of helpers/problem_helper.py:4-39 and short SDC runs modelled on tutorial/step_5/C and
    step_8/C (the latter with its cap of 10 GMRES iterations per solve)."""
    for tag, cls, pp in [
        ("advection1d_center_o2", advectionNd, dict(nvars=64, c=1.0, freq=2, stencil_type="center", order=2, bc="periodic", solver_type="GMRES", lintol=1e-12)),
        ("advection1d_upwind_o5", advectionNd, dict(nvars=64, c=0.7, freq=4, stencil_type="upwind", order=5, bc="periodic", solver_type="GMRES", lintol=1e-12)),
        ("advection1d_upwind_o1", advectionNd, dict(nvars=128, c=1.0, freq=2, stencil_type="upwind", order=1, bc="periodic", solver_type="GMRES", lintol=1e-10)),
        ("advection1d_forward_o3", advectionNd, dict(nvars=64, c=-1.0, freq=2, stencil_type="forward", order=3, bc="periodic", solver_type="GMRES", lintol=1e-12)),
        ("advection2d_center_o6", advectionNd, dict(nvars=(32, 32), c=0.1, freq=(2, 2), stencil_type="center", order=6, bc="periodic", solver_type="GMRES", lintol=1e-12)),
        ("advection2d_backward_o2", advectionNd, dict(nvars=(32, 32), c=1.0, freq=(2, 4), stencil_type="backward", order=2, bc="periodic", solver_type="GMRES", lintol=1e-12)),
        ("advection3d_upwind_o3", advectionNd, dict(nvars=(16, 16, 16), c=0.5, freq=(2, 2, 2), stencil_type="upwind", order=3, bc="periodic", solver_type="GMRES", lintol=1e-12)),
        ("advection2d_dirichlet_center_o4", advectionNd, dict(nvars=(31, 31), c=1.0, freq=(1, 2), stencil_type="center", order=4, bc="dirichlet-zero", solver_type="GMRES", lintol=1e-12)),
        ("advection1d_dirichlet_upwind_o4", advectionNd, dict(nvars=63, c=1.0, freq=3, stencil_type="upwind", order=4, bc="dirichlet-zero", solver_type="GMRES", lintol=1e-12)),
        ("heat2d_dirichlet_gmres", heatNd_unforced, dict(nvars=(31, 31), nu=0.1, freq=(2, 2), bc="dirichlet-zero", solver_type="GMRES", lintol=1e-12)),
        ("heat2d_periodic_o4_gmres", heatNd_forced, dict(nvars=(32, 32), nu=0.1, freq=(2, 2), bc="periodic", order=4, solver_type="GMRES", lintol=1e-12)),
        ("heat3d_dirichlet_o6_gmres", heatNd_unforced, dict(nvars=(15, 15, 15), nu=0.1, freq=(1, 1, 1), bc="dirichlet-zero", order=6, solver_type="GMRES", lintol=1e-12)),
        ("heat1d_dirichlet_gmres_capped", heatNd_unforced, dict(nvars=127, nu=0.7, freq=2, bc="dirichlet-zero", solver_type="GMRES", lintol=1e-13, liniter=33)),
    ]:
        P = cls(**pp)
        gen = np.random.default_rng(sum(map(ord, tag)))
        u = P.u_init
        u[:] = gen.standard_normal(u.shape)
        rhs = P.u_init
        rhs[:] = gen.standard_normal(u.shape)
        t, factor = 0.37, 0.0123
        f = P.eval_f(u, t)
        sol = P.solve_system(rhs, factor, u, t)
        spec = dict(problem=cls.__name__, problem_params=_jsonable(pp), t=t, factor=factor)
        save("op_" + tag, spec, u=np.asarray(u), rhs=np.asarray(rhs), f=np.asarray(f), sol=np.asarray(sol),
             gmres_iters=np.array(P.work_counters["GMRES"].niter), u_exact=np.asarray(P.u_exact(0.1)))
        print("   GMRES iterations:", P.work_counters["GMRES"].niter)
    # tutorial/step_5/C_advection_and_PFASST.py:17-35 on one level, with GMRES instead of the sparse direct solve
    spec = dict(problem="advectionNd", sweeper="generic_implicit",
                problem_params=dict(nvars=128, c=1, freq=4, order=4, bc="periodic", stencil_type="center",
                                    solver_type="GMRES", lintol=1e-12),
                sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU"),
                level_params=dict(dt=0.0625, restol=1e-9), step_params=dict(maxiter=50), t0=0.0, Tend=0.25, u0="exact")
    run_case("run_advection1d_gi_lu_gmres_128", spec)
    # tutorial/step_8/C_iteration_estimator.py:85-107: 2-D, order 6, at most 10 GMRES iterations per solve
    spec = dict(problem="advectionNd", sweeper="generic_implicit",
                problem_params=dict(nvars=[64, 64], c=0.1, freq=[2, 2], order=6, bc="periodic", stencil_type="center",
                                    solver_type="GMRES", liniter=10),
                sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU"),
                level_params=dict(dt=0.05, restol=1e-9), step_params=dict(maxiter=50), t0=0.0, Tend=0.1, u0="exact")
    run_case("run_advection2d_gi_lu_gmres10_64", spec)
    # upwind stencil with a diagonal QDelta
    spec = dict(problem="advectionNd", sweeper="generic_implicit",
                problem_params=dict(nvars=[32, 32, 32], c=1.0, freq=[2, 2, 2], order=3, bc="periodic",
                                    stencil_type="upwind", solver_type="GMRES", lintol=1e-12),
                sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="MIN-SR-NS"),
                level_params=dict(dt=0.01, restol=1e-9), step_params=dict(maxiter=50), t0=0.0, Tend=0.02, u0="exact")
    run_case("run_advection3d_gi_minsrns_gmres_32", spec)
    # heat equation on GMRES
    spec = dict(problem="heatNd_forced", sweeper="imex_1st_order",
                problem_params=dict(nvars=[63, 63], nu=0.1, freq=[4, 4], bc="dirichlet-zero", solver_type="GMRES",
                                    lintol=1e-12, liniter=10000),
                sweeper_params=dict(num_nodes=4, quad_type="RADAU-RIGHT", QI="LU"),
                level_params=dict(dt=0.1, restol=1e-10), step_params=dict(maxiter=50), t0=0.0, Tend=0.2, u0="exact")
    run_case("run_heat2d_imex_lu_gmres_63", spec)


def allencahn_semiimplicit_v2_fixtures():
    """allencahn_semiimplicit_v2 (AllenCahn_2D_FD.py:380-484): operator vectors and a short IMEX run."""
    rng = np.random.default_rng(4042)
    pp = dict(nvars=(32, 32), nu=2, eps=0.04, newton_maxiter=100, newton_tol=1e-9, lin_tol=1e-10, lin_maxiter=100, radius=0.25)
    P = allencahn_semiimplicit_v2(**pp)
    u = P.u_exact(0.0)
    u[:] = u + 0.01 * rng.standard_normal(u.shape)
    rhs = P.dtype_u(u)
    rhs[:] = u + 0.05 * rng.standard_normal(u.shape)
    f = P.eval_f(u, 0.0)
    sol = P.solve_system(rhs, 1e-3, u, 0.0)
    save("op_allencahn_semiimplicit_v2", dict(problem="allencahn_semiimplicit_v2", problem_params=_jsonable(pp), t=0.0, factor=1e-3),
         u=np.asarray(u), rhs=np.asarray(rhs), f=np.asarray(f), sol=np.asarray(sol), u_exact=np.asarray(P.u_exact(0.0)),
         newton_itercount=np.array(P.newton_itercount), newton=np.array(P.work_counters["newton"].niter),
         linear=np.array(P.work_counters["linear"].niter))
    spec = dict(problem="allencahn_semiimplicit_v2", sweeper="imex_1st_order",
                problem_params=dict(nvars=[64, 64], nu=2, eps=0.04, newton_maxiter=100, newton_tol=1e-9, lin_tol=1e-10,
                                    lin_maxiter=100, radius=0.25),
                sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU", initial_guess="zero"),
                level_params=dict(dt=1e-3, restol=1e-8), step_params=dict(maxiter=50), t0=0.0, Tend=2e-3, u0="exact")
    run_case("run_allencahn_semi_v2_imex_lu_64", spec)


def allencahn_multiimplicit_fixtures():
    """allencahn_multiimplicit / _v2 with the multi_implicit sweeper (AllenCahn_2D_FD.py:487-776,
    sweeper_classes/multi_implicit.py): operator vectors of both solves and the TOMS runs
    (projects/TOMS/AllenCahn_contracting_circle.py:36-62,114-124) at two sizes."""
    rng = np.random.default_rng(4043)
    pp = dict(nvars=(32, 32), nu=2, eps=0.04, newton_maxiter=100, newton_tol=1e-9, lin_tol=1e-10, lin_maxiter=100, radius=0.25)
    for name, cls in (("allencahn_multiimplicit", allencahn_multiimplicit), ("allencahn_multiimplicit_v2", allencahn_multiimplicit_v2)):
        P = cls(**pp)
        u = P.u_exact(0.0)
        u[:] = u + 0.01 * rng.standard_normal(u.shape)
        rhs = P.dtype_u(u)
        rhs[:] = u + 0.05 * rng.standard_normal(u.shape)
        f = P.eval_f(u, 0.0)
        sol1 = P.solve_system_1(rhs, 1e-3, u, 0.0)
        n1, l1 = int(P.newton_itercount), int(P.lin_itercount)
        sol2 = P.solve_system_2(rhs, 1e-3, u, 0.0)
        save("op_" + name, dict(problem=name, problem_params=_jsonable(pp), t=0.0, factor=1e-3), u=np.asarray(u),
             rhs=np.asarray(rhs), f=np.asarray(f), sol1=np.asarray(sol1), sol2=np.asarray(sol2),
             newton_after_1=np.array(n1), linear_after_1=np.array(l1), newton_itercount=np.array(int(P.newton_itercount)),
             lin_itercount=np.array(int(P.lin_itercount)))
        for n in (64, 128):
            spec = dict(problem=name, sweeper="multi_implicit",
                        problem_params=dict(nvars=[n, n], nu=2, eps=0.04, newton_maxiter=100, newton_tol=1e-9, lin_tol=1e-10,
                                            lin_maxiter=100, radius=0.25),
                        sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", Q1="LU", Q2="LU", initial_guess="zero"),
                        level_params=dict(dt=1e-3, restol=1e-8), step_params=dict(maxiter=50), t0=0.0, Tend=2e-3, u0="exact")
            tag = "multi" if name.endswith("implicit") else "multi_v2"
            run_case(f"run_allencahn_{tag}_lu_{n}", spec)


# ----------------------------------------------------------------------------------------------------------------
# full runs through the reference controller
# ----------------------------------------------------------------------------------------------------------------
def run_case(name, spec, store_uend=True):
    d = make_description(spec)
    c = controller_nonMPI(num_procs=1, controller_params={"logger_level": 40, "hook_class": [LogWork]}, description=d)
    P = c.MS[0].levels[0].prob
    u0 = initial_value(P, spec)
    uend, stats = c.run(u0=u0, t0=spec["t0"], Tend=spec["Tend"])
    niter = [int(v) for _, v in get_sorted(stats, type="niter", sortby="time")]
    times = [float(t) for t, _ in get_sorted(stats, type="niter", sortby="time")]
    res_hist = []
    for t in times:
        res_hist.append([float(v) for _, v in get_sorted(stats, time=t, type="residual_post_iteration", sortby="iter")])
    width = max(len(r) for r in res_hist)
    res = np.full((len(res_hist), width), np.nan)
    for i, r in enumerate(res_hist):
        res[i, : len(r)] = r
    arrays = {"niter": np.array(niter), "times": np.array(times), "residuals": res,
              "uend_maxabs": np.array(float(abs(uend)))}
    for key in P.work_counters:
        arrays["work_" + key] = np.array([int(v) for _, v in get_sorted(stats, type="work_" + key, sortby="time")])
    # (allencahn_semiimplicit_v2 inherits a u_exact(t > 0) that cannot digest its own imex right-hand side)
    if spec["u0"] == "exact" and spec["problem"] not in ("allencahn_fullyimplicit", "allencahn_semiimplicit_v2",
                                                         "allencahn_multiimplicit", "allencahn_multiimplicit_v2"):
        arrays["err_vs_exact"] = np.array(float(abs(P.u_exact(spec["Tend"]) - uend)))
    if store_uend:
        arrays["uend"] = np.asarray(uend)
    if hasattr(P, "newton_itercount"):
        arrays["newton_itercount"] = np.array(int(P.newton_itercount))
        arrays["newton_ncalls"] = np.array(int(P.newton_ncalls))
        arrays["lin_itercount"] = np.array(int(P.lin_itercount))
        arrays["lin_ncalls"] = np.array(int(P.lin_ncalls))
    save(name, spec, **arrays)
    return arrays


def full_runs():
    # config 1 (BASELINE.json configs[0]): 1-D heat, generic_implicit LU, direct solve, 20 steps
    spec = dict(problem="heatNd_unforced", sweeper="generic_implicit",
                problem_params=dict(nvars=1023, nu=0.1, freq=4, bc="dirichlet-zero"),
                sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU"),
                level_params=dict(dt=0.05, restol=1e-10), step_params=dict(maxiter=50),
                t0=0.0, Tend=1.0, u0="exact")
    a = run_case("run_heat1d_gi_lu_direct", spec)
    assert a["niter"].tolist() == [10, 10, 10, 9, 9, 9, 8, 8, 8, 7, 7, 6, 6, 6, 5, 5, 5, 4, 4, 4], a["niter"]

    # same with CG so that the device path (no sparse direct solver) has a 1-D pin
    spec_cg = json.loads(json.dumps(spec))
    spec_cg["problem_params"].update(solver_type="CG", lintol=1e-13, liniter=10000)
    run_case("run_heat1d_gi_lu_cg", spec_cg)

    # tutorial step_3 A: 1-D forced heat, IMEX IE, 12 iterations on every step
    spec = dict(problem="heatNd_forced", sweeper="imex_1st_order",
                problem_params=dict(nvars=1023, nu=0.1, freq=4, bc="dirichlet-zero"),
                sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT"),
                level_params=dict(dt=0.1, restol=1e-10), step_params=dict(maxiter=20),
                t0=0.1, Tend=0.9, u0="exact")
    a = run_case("run_heat1d_imex_ie_step3A", spec)
    assert all(n == 12 for n in a["niter"]), a["niter"]
    spec_cg = json.loads(json.dumps(spec))
    spec_cg["problem_params"].update(solver_type="CG", lintol=1e-13, liniter=10000)
    run_case("run_heat1d_imex_ie_cg", spec_cg)

    # config 2 scaled down: 2-D forced heat, IMEX LU M=4, CG
    for n in (63, 255):
        spec = dict(problem="heatNd_forced", sweeper="imex_1st_order",
                    problem_params=dict(nvars=[n, n], nu=0.1, freq=[4, 4], bc="dirichlet-zero", solver_type="CG",
                                        lintol=1e-12, liniter=10000),
                    sweeper_params=dict(num_nodes=4, quad_type="RADAU-RIGHT", QI="LU"),
                    level_params=dict(dt=0.1, restol=1e-10), step_params=dict(maxiter=50),
                    t0=0.0, Tend=0.1, u0="exact")
        run_case(f"run_heat2d_imex_lu_{n}", spec)

    # config 3 scaled down: 3-D unforced heat, generic_implicit, random initial data, two steps
    for qi in ("MIN-SR-NS", "LU", "IE"):
        spec = dict(problem="heatNd_unforced", sweeper="generic_implicit",
                    problem_params=dict(nvars=[31, 31, 31], nu=0.1, freq=[1, 1, 1], bc="dirichlet-zero",
                                        solver_type="CG", lintol=1e-12, liniter=10000),
                    sweeper_params=dict(num_nodes=4, quad_type="RADAU-RIGHT", QI=qi, initial_guess="spread"),
                    level_params=dict(dt=1e-3, restol=1e-8), step_params=dict(maxiter=50),
                    t0=0.0, Tend=2e-3, u0="random", seed=1234)
        a = run_case(f"run_heat3d_gi_{qi.lower().replace('-', '')}_31", spec)
        if qi == "MIN-SR-NS":
            assert a["niter"].tolist() == [7, 7] and a["work_CG"].tolist() == [142, 140], (a["niter"], a["work_CG"])
        if qi == "LU":
            assert a["niter"].tolist() == [9, 9] and a["work_CG"].tolist() == [204, 201], (a["niter"], a["work_CG"])
    # MIN-SR-FLEX as it is meant to be run: nsweeps = M (test_preconditioners.py:97-118)
    spec = dict(problem="heatNd_unforced", sweeper="generic_implicit",
                problem_params=dict(nvars=[31, 31, 31], nu=0.1, freq=[1, 1, 1], bc="dirichlet-zero",
                                    solver_type="CG", lintol=1e-12, liniter=10000),
                sweeper_params=dict(num_nodes=4, quad_type="RADAU-RIGHT", QI="MIN-SR-FLEX", initial_guess="spread"),
                level_params=dict(dt=1e-3, restol=1e-8, nsweeps=4), step_params=dict(maxiter=50),
                t0=0.0, Tend=2e-3, u0="random", seed=1234)
    run_case("run_heat3d_gi_minsrflex_31", spec)
    # fixed sweep count, as bench.py runs it (restol=-1, maxiter=K)
    spec = dict(problem="heatNd_unforced", sweeper="generic_implicit",
                problem_params=dict(nvars=[63, 63, 63], nu=0.1, freq=[1, 1, 1], bc="dirichlet-zero",
                                    solver_type="CG", lintol=1e-12, liniter=10000),
                sweeper_params=dict(num_nodes=4, quad_type="RADAU-RIGHT", QI="MIN-SR-NS", initial_guess="spread"),
                level_params=dict(dt=1e-3, restol=-1), step_params=dict(maxiter=4),
                t0=0.0, Tend=1e-3, u0="random", seed=1234)
    run_case("run_heat3d_gi_minsrns_63_K4", spec, store_uend=False)
    # periodic 2-D heat (even sizes, wrap-around stencil)
    spec = dict(problem="heatNd_unforced", sweeper="generic_implicit",
                problem_params=dict(nvars=[64, 64], nu=0.1, freq=[2, 2], bc="periodic", solver_type="CG",
                                    lintol=1e-12, liniter=10000),
                sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU"),
                level_params=dict(dt=0.01, restol=1e-10), step_params=dict(maxiter=50),
                t0=0.0, Tend=0.02, u0="random", seed=7)
    run_case("run_heat2d_gi_lu_periodic_64", spec)

    # config 4 scaled down: Allen-Cahn fully implicit, TOMS set-up (projects/TOMS/AllenCahn_contracting_circle.py:36-62)
    for n in (64, 128):
        spec = dict(problem="allencahn_fullyimplicit", sweeper="generic_implicit",
                    problem_params=dict(nvars=[n, n], nu=2, eps=0.04, newton_maxiter=100, newton_tol=1e-9,
                                        lin_tol=1e-10, lin_maxiter=100, radius=0.25),
                    sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU", initial_guess="zero"),
                    level_params=dict(dt=1e-3, restol=1e-8), step_params=dict(maxiter=50),
                    t0=0.0, Tend=2e-3, u0="exact")
        a = run_case(f"run_allencahn_gi_lu_{n}", spec)
        if n == 128:
            assert a["niter"].tolist() == [7, 7] and a["work_newton"].tolist() == [34, 36], (a["niter"], a["work_newton"])


# ----------------------------------------------------------------------------------------------------------------
# single-sweep dumps: every intermediate the sweeper API exposes, after predict + n sweeps on one Step
# ----------------------------------------------------------------------------------------------------------------
def sweep_dump(name, spec, nsweeps=2, with_tau=False):
    d = make_description(spec)
    S = Step(description=d)
    L = S.levels[0]
    P = L.prob
    L.status.time = spec["t0"]
    u0 = initial_value(P, spec)
    S.init_step(u0)
    L.sweep.predict()
    arrays = {"u0": np.asarray(u0)}
    M = L.sweep.coll.num_nodes
    if with_tau:
        rng = np.random.default_rng(99)
        for m in range(M):
            L.tau[m] = P.u_init
            L.tau[m][:] = 1e-3 * rng.standard_normal(L.tau[m].shape)
        arrays["tau"] = np.stack([np.asarray(t) for t in L.tau])
    arrays["f_pred"] = np.stack([np.asarray(f) for f in L.f])
    L.sweep.compute_residual()
    arrays["res_pred"] = np.array(L.status.residual)
    for k in range(1, nsweeps + 1):
        L.status.sweep = k
        L.sweep.updateVariableCoeffs(k)
        L.sweep.update_nodes()
        arrays[f"u_sweep{k}"] = np.stack([np.asarray(u) for u in L.u])
        arrays[f"f_sweep{k}"] = np.stack([np.asarray(f) for f in L.f])
        arrays[f"integrate_sweep{k}"] = np.stack([np.asarray(i) for i in L.sweep.integrate()])
        for rt in ("full_abs", "last_abs", "full_rel", "last_rel"):
            L.params.residual_type = rt
            L.sweep.compute_residual()
            arrays[f"res_{rt}_sweep{k}"] = np.array(L.status.residual)
        L.params.residual_type = spec["level_params"].get("residual_type", "full_abs")
        arrays[f"resvec_sweep{k}"] = np.stack([np.asarray(r) for r in L.residual])
    L.sweep.compute_end_point()
    arrays["uend"] = np.asarray(L.uend)
    arrays["QI"] = L.sweep.QI
    if hasattr(L.sweep, "QE"):
        arrays["QE"] = L.sweep.QE
    arrays["Qmat"] = L.sweep.coll.Qmat
    arrays["nodes"] = L.sweep.coll.nodes
    arrays["weights"] = L.sweep.coll.weights
    for key in P.work_counters:
        arrays["work_" + key] = np.array(P.work_counters[key].niter)
    save(name, spec, **arrays)


def sweep_dumps():
    base3d = dict(problem="heatNd_unforced", sweeper="generic_implicit",
                  problem_params=dict(nvars=[15, 15, 15], nu=0.1, freq=[1, 1, 1], bc="dirichlet-zero",
                                      solver_type="CG", lintol=1e-12, liniter=10000),
                  sweeper_params=dict(num_nodes=4, quad_type="RADAU-RIGHT", QI="MIN-SR-NS", initial_guess="spread"),
                  level_params=dict(dt=1e-3), step_params=dict(maxiter=50), t0=0.0, u0="random", seed=1234)
    sweep_dump("sweep_gi_minsrns_3d", base3d)
    base3d["problem_params"]["nvars"] = [9, 9, 9]  # keep the remaining 3-D dumps small
    s = json.loads(json.dumps(base3d)); s["sweeper_params"]["QI"] = "LU"
    sweep_dump("sweep_gi_lu_3d_tau", s, with_tau=True)
    s = json.loads(json.dumps(base3d)); s["sweeper_params"].update(QI="IE", quad_type="GAUSS", num_nodes=3)
    sweep_dump("sweep_gi_ie_gauss_3d", s)  # right end point not a node -> collocation update in compute_end_point
    s = json.loads(json.dumps(base3d)); s["sweeper_params"].update(QI="LU", quad_type="LOBATTO", num_nodes=3,
                                                                    do_coll_update=True, initial_guess="copy")
    sweep_dump("sweep_gi_lu_lobatto_3d", s)  # alpha == 0 on the first node: u = rhs branch (generic_implicit.py:93-94)
    s = json.loads(json.dumps(base3d)); s["sweeper_params"].update(QI="MIN-SR-FLEX"); s["level_params"]["nsweeps"] = 4
    sweep_dump("sweep_gi_minsrflex_3d", s, nsweeps=4)
    imex2d = dict(problem="heatNd_forced", sweeper="imex_1st_order",
                  problem_params=dict(nvars=[31, 31], nu=0.1, freq=[4, 4], bc="dirichlet-zero", solver_type="CG",
                                      lintol=1e-12, liniter=10000),
                  sweeper_params=dict(num_nodes=4, quad_type="RADAU-RIGHT", QI="LU"),
                  level_params=dict(dt=0.1), step_params=dict(maxiter=50), t0=0.3, u0="exact")
    sweep_dump("sweep_imex_lu_2d", imex2d)
    sweep_dump("sweep_imex_lu_2d_tau", imex2d, with_tau=True)
    s = json.loads(json.dumps(imex2d)); s["problem_params"].update(nvars=[9, 9, 9], freq=[2, 2, 2])
    s["sweeper_params"].update(QI="IE", num_nodes=3)
    sweep_dump("sweep_imex_ie_3d", s)
    s = json.loads(json.dumps(imex2d)); s["problem_params"].update(nvars=127, freq=2)
    s["sweeper_params"].update(num_nodes=3, QI="IE", quad_type="GAUSS")
    sweep_dump("sweep_imex_ie_gauss_1d", s)
    ac = dict(problem="allencahn_fullyimplicit", sweeper="generic_implicit",
              problem_params=dict(nvars=[32, 32], nu=2, eps=0.04, newton_maxiter=100, newton_tol=1e-9,
                                  lin_tol=1e-10, lin_maxiter=100, radius=0.25),
              sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU", initial_guess="zero"),
              level_params=dict(dt=1e-3), step_params=dict(maxiter=50), t0=0.0, u0="exact")
    sweep_dump("sweep_ac_lu", ac)


# ----------------------------------------------------------------------------------------------------------------
# operator-level vectors: eval_f / solve_system on seeded inputs
# ----------------------------------------------------------------------------------------------------------------
def operator_vectors():
    rng = np.random.default_rng(2024)
    for tag, cls, pp in [
        ("heat1d_dirichlet", heatNd_unforced, dict(nvars=63, nu=0.7, freq=2, bc="dirichlet-zero", solver_type="CG", lintol=1e-13)),
        ("heat2d_dirichlet", heatNd_unforced, dict(nvars=(31, 31), nu=0.1, freq=(2, 2), bc="dirichlet-zero", solver_type="CG", lintol=1e-13)),
        ("heat3d_dirichlet", heatNd_unforced, dict(nvars=(15, 15, 15), nu=0.1, freq=(1, 1, 1), bc="dirichlet-zero", solver_type="CG", lintol=1e-13)),
        ("heat1d_periodic", heatNd_unforced, dict(nvars=64, nu=0.3, freq=2, bc="periodic", solver_type="CG", lintol=1e-13)),
        ("heat2d_periodic", heatNd_unforced, dict(nvars=(32, 32), nu=0.1, freq=(2, 2), bc="periodic", solver_type="CG", lintol=1e-13)),
        ("heat3d_periodic", heatNd_unforced, dict(nvars=(16, 16, 16), nu=0.1, freq=(2, 2, 2), bc="periodic", solver_type="CG", lintol=1e-13)),
        ("heat2d_forced", heatNd_forced, dict(nvars=(31, 31), nu=0.1, freq=(4, 2), bc="dirichlet-zero", solver_type="CG", lintol=1e-13)),
        ("heat3d_forced", heatNd_forced, dict(nvars=(15, 15, 15), nu=0.1, freq=(1, 2, 3), bc="dirichlet-zero", solver_type="CG", lintol=1e-13)),
        # higher-order stencils (helpers/problem_helper.py:19-21; dirichlet: one-sided closure rows, :157-201)
        ("heat2d_periodic_o4", heatNd_unforced, dict(nvars=(32, 32), nu=0.1, freq=(2, 2), bc="periodic", order=4, solver_type="CG", lintol=1e-13)),
        ("heat2d_periodic_o8", heatNd_unforced, dict(nvars=(32, 32), nu=0.1, freq=(2, 2), bc="periodic", order=8, solver_type="CG", lintol=1e-13)),
        ("heat3d_periodic_o6", heatNd_unforced, dict(nvars=(16, 16, 16), nu=0.1, freq=(2, 2, 2), bc="periodic", order=6, solver_type="CG", lintol=1e-13)),
        ("heat1d_dirichlet_o8", heatNd_unforced, dict(nvars=63, nu=0.7, freq=2, bc="dirichlet-zero", order=8, solver_type="CG", lintol=1e-13)),
        ("heat2d_dirichlet_o4", heatNd_forced, dict(nvars=(31, 31), nu=0.1, freq=(4, 2), bc="dirichlet-zero", order=4, solver_type="CG", lintol=1e-13)),
        ("heat3d_dirichlet_o6", heatNd_unforced, dict(nvars=(15, 15, 15), nu=0.1, freq=(1, 1, 1), bc="dirichlet-zero", order=6, solver_type="CG", lintol=1e-13)),
    ]:
        P = cls(**pp)
        # (the higher-order cases were added later: own generator, so that the earlier fixtures regenerate unchanged)
        gen = np.random.default_rng(3000 + pp["order"]) if "order" in pp else rng
        u = P.u_init
        u[:] = gen.standard_normal(u.shape)
        rhs = P.u_init
        rhs[:] = gen.standard_normal(u.shape)
        t, factor = 0.37, 0.0123
        f = P.eval_f(u, t)
        sol = P.solve_system(rhs, factor, u, t)
        spec = dict(problem=cls.__name__, problem_params=_jsonable(pp), t=t, factor=factor)
        save("op_" + tag, spec, u=np.asarray(u), rhs=np.asarray(rhs), f=np.asarray(f), sol=np.asarray(sol),
             cg_iters=np.array(P.work_counters["CG"].niter), u_exact=np.asarray(P.u_exact(0.1)),
             A_row0=P.A[0].toarray().ravel()[:4] if P.A.shape[0] >= 4 else np.zeros(4),
             A_lastrow=P.A[-1].toarray().ravel()[-4:] if P.A.shape[0] >= 4 else np.zeros(4))
    pp = dict(nvars=(32, 32), nu=2, eps=0.04, newton_maxiter=100, newton_tol=1e-9, lin_tol=1e-10, lin_maxiter=100, radius=0.25)
    P = allencahn_fullyimplicit(**pp)
    u = P.u_exact(0.0)
    u[:] = u + 0.01 * rng.standard_normal(u.shape)
    rhs = P.dtype_u(u)
    f = P.eval_f(u, 0.0)
    sol = P.solve_system(rhs, 1e-3, u, 0.0)
    save("op_allencahn", dict(problem="allencahn_fullyimplicit", problem_params=_jsonable(pp), t=0.0, factor=1e-3),
         u=np.asarray(u), rhs=np.asarray(rhs), f=np.asarray(f), sol=np.asarray(sol), u_exact=np.asarray(P.u_exact(0.0)),
         newton=np.array(P.work_counters["newton"].niter), linear=np.array(P.work_counters["linear"].niter))


def allencahn_semiimplicit_fixtures():
    """allencahn_semiimplicit (AllenCahn_2D_FD.py:261-376): operator vectors and a short IMEX run (added in round 2)."""
    rng = np.random.default_rng(4041)
    pp = dict(nvars=(32, 32), nu=2, eps=0.04, newton_maxiter=100, newton_tol=1e-9, lin_tol=1e-10, lin_maxiter=100, radius=0.25)
    P = allencahn_semiimplicit(**pp)
    u = P.u_exact(0.0)
    u[:] = u + 0.01 * rng.standard_normal(u.shape)
    rhs = P.dtype_u(u)
    rhs[:] = rng.standard_normal(u.shape)
    f = P.eval_f(u, 0.0)
    sol = P.solve_system(rhs, 1e-3, u, 0.0)
    save("op_allencahn_semiimplicit", dict(problem="allencahn_semiimplicit", problem_params=_jsonable(pp), t=0.0, factor=1e-3),
         u=np.asarray(u), rhs=np.asarray(rhs), f=np.asarray(f), sol=np.asarray(sol), u_exact=np.asarray(P.u_exact(0.0)),
         linear=np.array(P.work_counters["linear"].niter))
    spec = dict(problem="allencahn_semiimplicit", sweeper="imex_1st_order",
                problem_params=dict(nvars=[64, 64], nu=2, eps=0.04, newton_maxiter=100, newton_tol=1e-9, lin_tol=1e-10,
                                    lin_maxiter=100, radius=0.25),
                sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU", initial_guess="zero"),
                level_params=dict(dt=1e-3, restol=1e-8), step_params=dict(maxiter=50), t0=0.0, Tend=2e-3, u0="exact")
    run_case("run_allencahn_semi_imex_lu_64", spec)


# ----------------------------------------------------------------------------------------------------------------
# space transfer: the reference's mesh_to_mesh (sparse Kronecker products) applied to seeded random fields
# ----------------------------------------------------------------------------------------------------------------
def transfer_vectors():
    rng = np.random.default_rng(77)
    cases = [
        ("1d_dirichlet", heatNd_unforced, dict(nu=0.1, freq=2, bc="dirichlet-zero"), 63, 31, dict(rorder=2, iorder=6)),
        ("2d_dirichlet", heatNd_forced, dict(nu=0.1, freq=(2, 2), bc="dirichlet-zero"), (31, 31), (15, 15), dict(rorder=2, iorder=6)),
        ("2d_dirichlet_o4", heatNd_unforced, dict(nu=0.1, freq=(2, 2), bc="dirichlet-zero"), (31, 31), (15, 15), dict(rorder=2, iorder=4)),
        ("3d_dirichlet", heatNd_unforced, dict(nu=0.1, freq=(1, 1, 1), bc="dirichlet-zero"), (15, 15, 15), (7, 7, 7), dict(rorder=2, iorder=2)),
        ("2d_periodic", heatNd_unforced, dict(nu=0.1, freq=(2, 2), bc="periodic"), (32, 32), (16, 16), dict(rorder=2, iorder=6, periodic=True)),
        ("1d_periodic", heatNd_unforced, dict(nu=0.1, freq=2, bc="periodic"), 64, 32, dict(rorder=2, iorder=4, periodic=True)),
    ]
    for tag, cls, pp, nf, nc, tp in cases:
        Pf, Pc = cls(nvars=nf, **pp), cls(nvars=nc, **pp)
        T = mesh_to_mesh(Pf, Pc, tp)
        F, G = Pf.u_init, Pc.u_init
        F[:] = rng.standard_normal(F.shape)
        G[:] = rng.standard_normal(G.shape)
        out = dict(F=np.asarray(F), G=np.asarray(G), RF=np.asarray(T.restrict(F)), PG=np.asarray(T.prolong(G)))
        if cls is heatNd_forced:  # multi-component right-hand sides go through component by component
            Ff, Gf = Pf.f_init, Pc.f_init
            Ff[:] = rng.standard_normal(Ff.shape)
            Gf[:] = rng.standard_normal(Gf.shape)
            out.update(Ff=np.asarray(Ff), Gf=np.asarray(Gf), RFf=np.asarray(T.restrict(Ff)), PGf=np.asarray(T.prolong(Gf)))
        spec = dict(problem=cls.__name__, problem_params=_jsonable(pp), nvars_fine=_jsonable(dict(n=nf))["n"],
                    nvars_coarse=_jsonable(dict(n=nc))["n"], transfer_params=tp)
        save("transfer_" + tag, spec, **out)


# ----------------------------------------------------------------------------------------------------------------
# PFASST (config 5 scaled down) through the reference's virtual-parallel controller
# ----------------------------------------------------------------------------------------------------------------
def pfasst_runs():
    # the reference's own known answer first (tutorial step_8 A: 7 iterations on all 8 steps)
    from pySDC.tutorial.step_6.A_run_non_MPI_controller import set_parameters_ml
    d, cp, t0, Tend = set_parameters_ml()
    c = controller_nonMPI(num_procs=8, controller_params=cp, description=d)
    P = c.MS[0].levels[0].prob
    uend, stats = c.run(u0=P.u_exact(t0), t0=t0, Tend=Tend)
    niter = [int(v) for _, v in get_sorted(stats, type="niter", sortby="time")]
    assert niter == [7] * 8 and abs(P.u_exact(Tend) - uend) < 6.1555e-05, niter
    save("pfasst_step8A_heat1d", dict(note="tutorial/step_6/A set_parameters_ml, num_procs=8", t0=t0, Tend=Tend),
         niter=np.array(niter), uend=np.asarray(uend), err=np.array(float(abs(P.u_exact(Tend) - uend))))

    for n, nprocs in ((63, 4), (127, 8)):
        spec = dict(problem="heatNd_forced", sweeper="imex_1st_order",
                    problem_params=dict(nvars=[[n, n], [n // 2, n // 2]], nu=0.1, freq=[4, 4], bc="dirichlet-zero",
                                        solver_type="CG", lintol=1e-12, liniter=10000),
                    sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU"),
                    level_params=dict(dt=0.25, restol=1e-10), step_params=dict(maxiter=50),
                    space_transfer_params=dict(rorder=2, iorder=6),
                    controller_params=dict(logger_level=40, predict_type="pfasst_burnin"),
                    num_procs=nprocs, t0=0.0, Tend=0.25 * nprocs, u0="exact")
        pp = dict(spec["problem_params"])
        pp["nvars"] = [tuple(v) for v in pp["nvars"]]
        pp["freq"] = tuple(pp["freq"])
        d = dict(problem_class=heatNd_forced, problem_params=pp, sweeper_class=imex_1st_order,
                 sweeper_params=dict(spec["sweeper_params"]), level_params=dict(spec["level_params"]),
                 step_params=dict(spec["step_params"]), space_transfer_class=mesh_to_mesh,
                 space_transfer_params=dict(spec["space_transfer_params"]))
        c = controller_nonMPI(num_procs=nprocs, controller_params=dict(spec["controller_params"], hook_class=[LogWork]),
                              description=d)
        P = c.MS[0].levels[0].prob
        uend, stats = c.run(u0=P.u_exact(0.0), t0=0.0, Tend=spec["Tend"])
        niter = [int(v) for _, v in get_sorted(stats, type="niter", sortby="time")]
        save(f"pfasst_heat2d_imex_{n}_p{nprocs}", spec, niter=np.array(niter), uend=np.asarray(uend),
             err=np.array(float(abs(P.u_exact(spec["Tend"]) - uend))))
        print("  PFASST", n, nprocs, niter)


def pfasst_config5():
    """BASELINE config 5 at FULL size (1023^2 / 511^2, 8 slices) through the reference's virtual-parallel controller.
    Slow (minutes of CPU); the fixture keeps the iteration counts, the error against the exact solution, the max-norm
    and a 64x64 subsample of uend (every 16th point) instead of the 8 MB field."""
    import time
    n, nprocs = 1023, 8
    spec = dict(problem="heatNd_forced", sweeper="imex_1st_order",
                problem_params=dict(nvars=[[n, n], [n // 2, n // 2]], nu=0.1, freq=[4, 4], bc="dirichlet-zero",
                                    solver_type="CG", lintol=1e-12, liniter=10000),
                sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU"),
                level_params=dict(dt=0.25, restol=1e-10), step_params=dict(maxiter=50),
                space_transfer_params=dict(rorder=2, iorder=6),
                controller_params=dict(logger_level=40, predict_type="pfasst_burnin"),
                num_procs=nprocs, t0=0.0, Tend=0.25 * nprocs, u0="exact", subsample=16)
    pp = dict(spec["problem_params"])
    pp["nvars"] = [tuple(v) for v in pp["nvars"]]
    pp["freq"] = tuple(pp["freq"])
    d = dict(problem_class=heatNd_forced, problem_params=pp, sweeper_class=imex_1st_order,
             sweeper_params=dict(spec["sweeper_params"]), level_params=dict(spec["level_params"]),
             step_params=dict(spec["step_params"]), space_transfer_class=mesh_to_mesh,
             space_transfer_params=dict(spec["space_transfer_params"]))
    t_start = time.perf_counter()
    c = controller_nonMPI(num_procs=nprocs, controller_params=dict(spec["controller_params"], hook_class=[LogWork]),
                          description=d)
    P = c.MS[0].levels[0].prob
    uend, stats = c.run(u0=P.u_exact(0.0), t0=0.0, Tend=spec["Tend"])
    wall = time.perf_counter() - t_start
    niter = [int(v) for _, v in get_sorted(stats, type="niter", sortby="time")]
    cg = [int(v) for _, v in get_sorted(stats, type="work_CG", sortby="time")]
    # residual after every iteration of every slice (rows = slices in time order, NaN-padded): the stopping decisions
    # of this run sit close to restol, the histories document how close
    hist = {}
    for k, v in stats.items():
        if k.type == "residual_post_iteration":
            hist.setdefault(round(k.time, 10), {})[k.iter] = float(v)
    res = np.full((len(hist), max(niter)), np.nan)
    for i, t in enumerate(sorted(hist)):
        for it, v in hist[t].items():
            res[i, it - 1] = v
    save(f"pfasst_config5_{n}_p{nprocs}", spec, niter=np.array(niter), uend_sub=np.asarray(uend)[::16, ::16].copy(),
         residuals=res,
         uend_maxnorm=np.array(float(abs(uend))), err=np.array(float(abs(P.u_exact(spec["Tend"]) - uend))),
         work_cg=np.array(cg), wall_seconds=np.array(wall))
    print("  PFASST config 5", niter, "wall", wall)


FAMILIES = {"gmres": gmres_fixtures, "allencahn_multi": allencahn_multiimplicit_fixtures, "allencahn_semi_v2": allencahn_semiimplicit_v2_fixtures, "allencahn_semi": allencahn_semiimplicit_fixtures, "runs": full_runs, "sweeps": sweep_dumps, "ops": operator_vectors, "transfer": transfer_vectors,
            "pfasst": pfasst_runs,
            "pfasst_config5": pfasst_config5}

if __name__ == "__main__":
    which = sys.argv[1:] or [k for k in FAMILIES if k != "pfasst_config5"]  # the full-size run is opt-in (slow)
    for w in which:
        FAMILIES[w]()
