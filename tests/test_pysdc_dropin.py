"""Drop-in check inside the UNMODIFIED reference: pySDC's own controller_nonMPI, Step, Level, hooks and convergence
controllers drive the classes of pysdc_b200.pysdc_plugin selected purely through the description dict.

Two variants of every test: ``numpy`` — the kernel library replaced by the numpy test double (CPU suite, host logic
only) — and ``cuda`` (``-m gpu``) — the reference's controller on the REAL CUDA kernels through the C ABI.  The
reference is imported from /root/reference in the build container and from the shipped copy ``oracle/_ref`` on the GPU
box (oracle/build_ref.py; verified against its manifest)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, reference_paths

REF_PATHS = reference_paths()
pytestmark = pytest.mark.skipif(REF_PATHS is None, reason="reference tree not present (no /root/reference, no oracle/_ref)")


@pytest.fixture(params=["numpy", pytest.param("cuda", marks=pytest.mark.gpu)])
def plugin(request):
    for p in reversed(REF_PATHS):
        if p not in sys.path:
            sys.path.insert(0, p)
    from pysdc_b200 import backend

    old = backend._backend
    if request.param == "cuda":
        backend.set_backend(backend.CudaBackend())
    else:
        from fake_backend import NumpyBackend

        backend.set_backend(NumpyBackend())
    from pysdc_b200 import pysdc_plugin

    yield pysdc_plugin
    backend.set_backend(old)


def test_shipped_reference_is_unmodified():
    """oracle/_ref (when present) is byte-identical to what oracle/build_ref.py copied from the reference tree."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref

    if not os.path.isfile(build_ref.MANIFEST):
        pytest.skip("oracle/_ref not built")
    assert build_ref.verify()


@pytest.mark.parametrize("name", ["run_heat3d_gi_minsrns_31", "run_heat3d_gi_lu_31", "run_heat2d_imex_lu_63",
                                  "run_heat1d_imex_ie_step3A", "run_allencahn_gi_lu_64", "run_heat3d_gi_minsrflex_31",
                                  "run_allencahn_semi_imex_lu_64", "run_allencahn_semi_v2_imex_lu_64",
                                  "run_advection2d_gi_lu_gmres10_64", "run_advection3d_gi_minsrns_gmres_32",
                                  "run_heat2d_imex_lu_gmres_63", "run_allencahn_multi_lu_64",
                                  "run_allencahn_multi_v2_lu_64"])
def test_reference_controller_drives_plugin_classes(plugin, name):
    from pySDC.core.sweeper import Sweeper
    from pySDC.helpers.stats_helper import get_sorted
    from pySDC.implementations.controller_classes.controller_nonMPI import controller_nonMPI
    from pySDC.implementations.hooks.log_work import LogWork

    spec, g = load_golden(name)
    pp = dict(spec["problem_params"])
    for k in ("nvars", "freq"):
        if isinstance(pp.get(k), list):
            pp[k] = tuple(pp[k])
    description = dict(problem_class=getattr(plugin, spec["problem"]), problem_params=pp,
                       sweeper_class=getattr(plugin, spec["sweeper"]), sweeper_params=dict(spec["sweeper_params"]),
                       level_params=dict(spec["level_params"]), step_params=dict(spec["step_params"]))
    c = controller_nonMPI(num_procs=1, controller_params={"logger_level": 40, "hook_class": [LogWork]},
                          description=description)
    L = c.MS[0].levels[0]
    assert isinstance(L.sweep, Sweeper)
    P = L.prob
    if spec["u0"] == "exact":
        u0 = P.u_exact(spec["t0"])
    else:
        u0 = P.u_init
        u0[:] = np.random.default_rng(spec["seed"]).standard_normal(P.nvars)
    uend, stats = c.run(u0=u0, t0=spec["t0"], Tend=spec["Tend"])
    niter = [int(v) for _, v in get_sorted(stats, type="niter", sortby="time")]
    assert niter == g["niter"].tolist()
    assert np.max(np.abs(uend.get() - g["uend"])) <= 1e-10 * max(abs(u0), float(g["uend_maxabs"]))
    for key in P.work_counters:
        got = [int(v) for _, v in get_sorted(stats, type="work_" + key, sortby="time")]
        want = g["work_" + key].tolist()
        band = 0.05 if key == "GMRES" else 0.02  # (restarted GMRES: see tests/test_host_cpu.py::test_run)
        assert np.all(np.abs(np.array(got) - np.array(want)) <= np.maximum(np.ceil(band * np.array(want)), 1)), (key, got, want)


def test_reference_pfasst_controller_drives_plugin_classes(plugin):
    """PFASST through the reference's OWN virtual-parallel controller, base transfer and convergence controllers, with
    the device problem / sweeper / space-transfer classes plugged in: same iteration counts and end value as the
    reference's classes gave (fixture pfasst_heat2d_imex_63_p4)."""
    from pySDC.helpers.stats_helper import get_sorted
    from pySDC.implementations.controller_classes.controller_nonMPI import controller_nonMPI

    from pysdc_b200.transfer import mesh_to_mesh

    spec, g = load_golden("pfasst_heat2d_imex_63_p4")
    pp = dict(spec["problem_params"])
    pp["nvars"] = [tuple(v) for v in pp["nvars"]]
    pp["freq"] = tuple(pp["freq"])
    d = dict(problem_class=getattr(plugin, spec["problem"]), problem_params=pp,
             sweeper_class=getattr(plugin, spec["sweeper"]), sweeper_params=dict(spec["sweeper_params"]),
             level_params=dict(spec["level_params"]), step_params=dict(spec["step_params"]),
             space_transfer_class=mesh_to_mesh, space_transfer_params=dict(spec["space_transfer_params"]))
    c = controller_nonMPI(num_procs=spec["num_procs"], controller_params=dict(spec["controller_params"]), description=d)
    P = c.MS[0].levels[0].prob
    uend, stats = c.run(u0=P.u_exact(0.0), t0=0.0, Tend=spec["Tend"])
    niter = [int(v) for _, v in get_sorted(stats, type="niter", sortby="time")]
    assert niter == g["niter"].tolist()
    assert np.max(np.abs(uend.get() - g["uend"])) / np.max(np.abs(g["uend"])) < 1e-10


def test_reference_hooks_work_on_device_fields(plugin):
    """The reference's logging hooks (solution, work, iteration counts, step size, timings, errors against u_exact) read
    L.uend / L.u / L.status / work_counters of the plug-in classes without modification."""
    from pySDC.helpers.stats_helper import get_sorted
    from pySDC.implementations.controller_classes.controller_nonMPI import controller_nonMPI
    from pySDC.implementations.hooks.log_errors import LogGlobalErrorPostRun, LogLocalErrorPostStep
    from pySDC.implementations.hooks.log_solution import LogSolution
    from pySDC.implementations.hooks.log_step_size import LogStepSize
    from pySDC.implementations.hooks.log_timings import CPUTimings
    from pySDC.implementations.hooks.log_work import LogSDCIterations, LogWork

    hooks = [LogSolution, LogWork, LogSDCIterations, LogStepSize, CPUTimings, LogGlobalErrorPostRun,
             LogLocalErrorPostStep]
    d = dict(problem_class=plugin.heatNd_unforced,
             problem_params=dict(nvars=(15, 15), nu=0.1, freq=(2, 2), bc="dirichlet-zero", solver_type="CG", lintol=1e-12,
                                 liniter=1000),
             sweeper_class=plugin.generic_implicit, sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU"),
             level_params=dict(dt=0.01, restol=1e-9), step_params=dict(maxiter=20))
    c = controller_nonMPI(num_procs=1, controller_params=dict(logger_level=40, hook_class=hooks), description=d)
    P = c.MS[0].levels[0].prob
    uend, stats = c.run(u0=P.u_exact(0.0), t0=0.0, Tend=0.03)
    sols = get_sorted(stats, type="u", sortby="time")
    assert len(sols) == 3 and type(sols[-1][1]) is plugin.mesh
    assert np.array_equal(sols[-1][1].get(), uend.get())
    assert sum(v for _, v in get_sorted(stats, type="work_CG")) == P.work_counters["CG"].niter
    assert [v for _, v in get_sorted(stats, type="k")] == [v for _, v in get_sorted(stats, type="niter")]
    err = get_sorted(stats, type="e_global_post_run")[-1][1]
    assert err == pytest.approx(abs(P.u_exact(0.03) - uend)) and err < 1e-3


def test_reference_LogToFile_hook_writes_device_fields(plugin, tmp_path):
    """the reference's own LogToFile hook (hooks/log_solution.py:207-282) asks the plug-in problem for
    getOutputFile / processSolutionForOutput; the file it writes is read back with the reference's FieldsIO, and the
    hook's resume path (re-opening the file with FieldsIO.fromFile and appending staged device fields) works."""
    from pySDC.helpers.fieldsIO import FieldsIO
    from pySDC.implementations.controller_classes.controller_nonMPI import controller_nonMPI
    from pySDC.implementations.hooks.log_solution import LogToFile

    class Log(LogToFile):
        filename = str(tmp_path / "heat.pySDC")

    d = dict(problem_class=plugin.heatNd_forced,
             problem_params=dict(nvars=(15, 15), nu=0.1, freq=(2, 2), bc="dirichlet-zero", solver_type="CG", lintol=1e-12,
                                 liniter=1000),
             sweeper_class=plugin.imex_1st_order, sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU"),
             level_params=dict(dt=0.01, restol=1e-9), step_params=dict(maxiter=20))
    c = controller_nonMPI(num_procs=1, controller_params=dict(logger_level=40, hook_class=[Log]), description=d)
    P = c.MS[0].levels[0].prob
    u0 = P.u_exact(0.0)
    umid, _ = c.run(u0=u0, t0=0.0, Tend=0.02)
    c.MS[0].levels[0].prob  # noqa: B018
    del c  # (drops the hook and with it the file object: a pending device field is flushed)
    import gc

    gc.collect()
    f = FieldsIO.fromFile(Log.filename)
    assert f.times == pytest.approx([0.0, 0.01, 0.02]) and f.gridSizes == [15, 15]
    assert np.array_equal(f.readField(0)[1][0], u0.get()) and np.array_equal(f.readField(-1)[1][0], umid.get())
    c = controller_nonMPI(num_procs=1, controller_params=dict(logger_level=40, hook_class=[Log]), description=d)
    uend, _ = c.run(u0=umid, t0=0.02, Tend=0.04)
    assert Log.load(-1)["t"] == pytest.approx(0.04) and np.array_equal(Log.load(-1)["u"][0], uend.get())


@pytest.mark.parametrize("name", ["run_allencahn_multi_lu_64", "run_allencahn_multi_v2_lu_64", "run_heat2d_imex_lu_63",
                                  "run_heat3d_gi_lu_31"])
def test_reference_sweepers_drive_plugin_problem_classes(plugin, name):
    """Only the PROBLEM class is swapped: the reference's unmodified sweepers (multi_implicit, imex_1st_order,
    generic_implicit) run their numpy-style arithmetic on the device datatypes (``dt * Q * f.comp1``, ``u - Q2int``,
    rebinding ``L.u[m+1] = P.solve_system_1(...)``) and call ``solve_system[_1/_2]`` / ``eval_f`` of the plug-in classes."""
    import importlib

    from pySDC.helpers.stats_helper import get_sorted
    from pySDC.implementations.controller_classes.controller_nonMPI import controller_nonMPI

    spec, g = load_golden(name)
    pp = dict(spec["problem_params"])
    for k in ("nvars", "freq"):
        if isinstance(pp.get(k), list):
            pp[k] = tuple(pp[k])
    sweeper = getattr(importlib.import_module("pySDC.implementations.sweeper_classes." + spec["sweeper"]), spec["sweeper"])
    d = dict(problem_class=getattr(plugin, spec["problem"]), problem_params=pp, sweeper_class=sweeper,
             sweeper_params=dict(spec["sweeper_params"]), level_params=dict(spec["level_params"]),
             step_params=dict(spec["step_params"]))
    c = controller_nonMPI(num_procs=1, controller_params={"logger_level": 40}, description=d)
    P = c.MS[0].levels[0].prob
    if spec["u0"] == "exact":
        u0 = P.u_exact(spec["t0"])
    else:
        u0 = P.u_init
        u0[:] = np.random.default_rng(spec["seed"]).standard_normal(P.nvars)
    uend, stats = c.run(u0=u0, t0=spec["t0"], Tend=spec["Tend"])
    assert [int(v) for _, v in get_sorted(stats, type="niter", sortby="time")] == g["niter"].tolist()
    assert np.max(np.abs(uend.get() - g["uend"])) <= 1e-10 * max(float(abs(u0)), float(g["uend_maxabs"]))
    if "newton_itercount" in g:
        assert P.newton_itercount == int(g["newton_itercount"]) and P.newton_ncalls == int(g["newton_ncalls"])


def test_reference_adaptivity_runs_on_plugin_classes(plugin):
    """The reference's step-size controller (convergence_controller_classes/adaptivity.py with EstimateEmbeddedError,
    StepSizeLimiter-free, restarts through BasicRestarting) on the plug-in classes against the same run on the
    reference's own classes: same accepted steps, same restarts, same step sizes.  The embedded estimate is a 1e-6-sized
    difference of two O(1) iterates computed with lintol = 1e-12 solves, so step sizes agree to ~1e-6 relative and the end
    values to the accuracy that step-size jitter allows (the run's own tolerance is e_tol = 1e-6)."""
    from pySDC.helpers.stats_helper import get_sorted
    from pySDC.implementations.controller_classes.controller_nonMPI import controller_nonMPI
    from pySDC.implementations.convergence_controller_classes.adaptivity import Adaptivity
    from pySDC.implementations.problem_classes.HeatEquation_ND_FD import heatNd_forced
    from pySDC.implementations.sweeper_classes.imex_1st_order import imex_1st_order

    def run(problem_class, sweeper_class):
        d = dict(problem_class=problem_class,
                 problem_params=dict(nvars=(63, 63), nu=0.1, freq=(2, 2), bc="dirichlet-zero", solver_type="CG", lintol=1e-12,
                                     liniter=10000),
                 sweeper_class=sweeper_class, sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU"),
                 level_params=dict(dt=0.05, restol=-1), step_params=dict(maxiter=3),
                 convergence_controllers={Adaptivity: dict(e_tol=1e-6)})
        c = controller_nonMPI(num_procs=1, controller_params=dict(logger_level=40, mssdc_jac=False), description=d)
        P = c.MS[0].levels[0].prob
        uend, stats = c.run(u0=P.u_exact(0.0), t0=0.0, Tend=0.25)
        dts = [v for _, v in get_sorted(stats, type="dt", recomputed=False)]
        restarts = sum(v for _, v in get_sorted(stats, type="restart"))
        return np.asarray(uend.get() if hasattr(uend, "get") else uend), np.array(dts), restarts

    u_ref, dt_ref, restarts_ref = run(heatNd_forced, imex_1st_order)
    u_dev, dt_dev, restarts_dev = run(plugin.heatNd_forced, plugin.imex_1st_order)
    assert len(dt_dev) == len(dt_ref) > 5 and restarts_dev == restarts_ref
    np.testing.assert_allclose(dt_dev, dt_ref, rtol=2e-5)
    assert np.max(np.abs(u_dev - u_ref)) < 1e-7
