"""Node-parallel ("parallel across the method") sweepers: one collocation node per rank / GPU.

Expected values are fixtures of the reference's OWN ``generic_implicit_MPI`` / ``imex_1st_order_MPI`` sweepers
(sweeper_classes/generic_implicit_MPI.py, imex_1st_order_MPI.py), unmodified, run with one process per node by
oracle/make_golden_node_parallel.py (which also checks them against the reference's serial sweepers).

Three ways through the code, each as a ``numpy`` variant (gloo + the numpy test double, CPU suite) and a ``cuda`` variant
(``-m gpu``: the real kernels; NCCL when the box has a GPU per node, otherwise the ranks share the device and the
collectives are staged through gloo):

* ``own``        pysdc_b200's sweepers + problem classes under pysdc_b200's stand-alone controller;
* ``plugin``     the same classes bound to pySDC's bases under the reference's own ``controller_nonMPI``;
* ``reference``  the reference's unmodified ``*_MPI`` sweepers (M ``comm.Reduce`` of full fields per quadrature) on the
                 mpi4py facade, driving the plug-in PROBLEM classes — device fields through ``Reduce / Allreduce / Bcast``.
"""
import json
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, free_port, load_golden, reference_paths

REF_PATHS = reference_paths()
CASES = ["nodepar_heat3d_gi_minsrns_31", "nodepar_heat2d_gi_minsrs_collupdate_63", "nodepar_heat2d_imex_minsrs_pic_63"]


def _worker(rank, world, port, name, out_dir, kind, mode, ref_paths):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SDCB200_CHECK_TAGS="1")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    if mode != "own":
        for p in reversed(ref_paths):
            sys.path.insert(0, p)
        import pysdc_b200.mpi_facade

        sys.path.insert(0, pysdc_b200.mpi_facade.PATH)
    transport = "gloo"
    if kind == "cuda":
        import torch

        transport = "nccl" if torch.cuda.device_count() >= world else "gloo"
        torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group(transport, rank=rank, world_size=world)
    try:
        from pysdc_b200 import backend

        if kind == "cuda":
            backend.set_backend(backend.CudaBackend())
        else:
            from fake_backend import NumpyBackend

            backend.set_backend(NumpyBackend())
        spec, _ = load_golden(name)
        pp = dict(spec["problem_params"])
        pp["nvars"], pp["freq"] = tuple(pp["nvars"]), tuple(pp["freq"])
        sp = dict(spec["sweeper_params"])
        if mode == "own":
            from pysdc_b200 import problems, sweepers
            from pysdc_b200.controller import LogWork, controller_nonMPI
            from pysdc_b200.parallel import TorchComm
            from pysdc_b200.stats import get_sorted

            prob, sweep, sp["comm"] = getattr(problems, spec["problem"]), getattr(sweepers, spec["sweeper"] + "_MPI"), TorchComm()
        else:
            from mpi4py import MPI  # the facade
            from pySDC.helpers.stats_helper import get_sorted
            from pySDC.implementations.controller_classes.controller_nonMPI import controller_nonMPI
            from pySDC.implementations.hooks.log_work import LogWork

            from pysdc_b200 import pysdc_plugin as plugin

            prob, sp["comm"] = getattr(plugin, spec["problem"]), MPI.COMM_WORLD
            if mode == "plugin":
                sweep = getattr(plugin, spec["sweeper"] + "_MPI")
            else:
                import importlib

                mod = importlib.import_module(f"pySDC.implementations.sweeper_classes.{spec['sweeper']}_MPI")
                sweep = getattr(mod, spec["sweeper"] + "_MPI")
        d = dict(problem_class=prob, problem_params=pp, sweeper_class=sweep, sweeper_params=sp,
                 level_params=dict(spec["level_params"]), step_params=dict(spec["step_params"]))
        c = controller_nonMPI(num_procs=1, controller_params={"logger_level": 40, "hook_class": [LogWork]}, description=d)
        P = c.MS[0].levels[0].prob
        if spec["u0"] == "exact":
            u0 = P.u_exact(spec["t0"])
        else:
            u0 = P.u_init
            u0[:] = np.random.default_rng(spec["seed"]).standard_normal(P.nvars)
        launches0 = backend.get_backend().launches
        uend, stats = c.run(u0=u0, t0=spec["t0"], Tend=spec["Tend"])
        its = get_sorted(stats, type="niter", sortby="time")
        res = [[float(v) for _, v in get_sorted(stats, time=t, type="residual_post_iteration", sortby="iter")] for t, _ in its]
        out = dict(niter=[int(v) for _, v in its], residuals=res,
                   work_CG=[int(v) for _, v in get_sorted(stats, type="work_CG", sortby="time")],
                   launches=backend.get_backend().launches - launches0)
        with open(os.path.join(out_dir, f"out_{rank}.json"), "w") as f:
            json.dump(out, f)
        np.save(os.path.join(out_dir, f"uend_{rank}.npy"), uend.get())
        if mode == "own" and rank == 0 and not sp.get("do_coll_update", False):
            # the gathered quadrature sums over the nodes in the serial sweeper's order: same bits as the serial
            # (node-batched) diagonal-QDelta sweep on one device
            sp.pop("comm")
            d.update(sweeper_class=getattr(sweepers, spec["sweeper"]), sweeper_params=sp)
            c = controller_nonMPI(num_procs=1, controller_params={"logger_level": 40}, description=d)
            serial, _ = c.run(u0=u0, t0=spec["t0"], Tend=spec["Tend"])
            assert np.array_equal(serial.get(), uend.get())
    finally:
        dist.destroy_process_group()


def _check(tmp_path, name, kind, mode):
    spec, g = load_golden(name)
    world = spec["sweeper_params"]["num_nodes"]
    mp.spawn(_worker, args=(world, free_port(), name, str(tmp_path), kind, mode, REF_PATHS), nprocs=world, join=True)
    outs = [json.load(open(os.path.join(tmp_path, f"out_{r}.json"))) for r in range(world)]
    uends = [np.load(os.path.join(tmp_path, f"uend_{r}.npy")) for r in range(world)]
    for r in range(world):
        assert outs[r]["niter"] == g["niter"].tolist(), (r, outs[r]["niter"])
        assert np.array_equal(uends[r], uends[0])  # the broadcast / the all-gathered quadrature: same bits on every rank
        assert np.max(np.abs(uends[r] - g["uend"])) / np.max(np.abs(g["uend"])) < 1e-10
        assert outs[r]["launches"] > 0
        for got, want in zip(outs[r]["residuals"], g["residuals"]):
            want = want[~np.isnan(want)]
            assert len(got) == len(want)
            # 1e-6 relative above the solver noise floor, like the serial runs (parity_cases.check_run)
            np.testing.assert_allclose(got, want, rtol=1e-6, atol=2e-11 * max(1.0, float(g["uend_maxabs"])))
        # every rank solves its own node: CG work per rank and step as the reference's ranks did it (the stopping test
        # ||r|| < 1e-12 ||b|| of the late sweeps, whose initial guess is almost the solution, is decided by rounding: +-2 %,
        # or one iteration per solve of the step)
        want = g["work_CG_per_rank"][r]
        slack = np.maximum(np.ceil(0.02 * want), g["niter"])
        assert np.all(np.abs(np.array(outs[r]["work_CG"]) - want) <= slack), (r, outs[r]["work_CG"], want)


def _variants(cuda_cases, numpy_cases=CASES):
    """(name, kind) pairs: ``numpy_cases`` on the numpy double (CPU suite), ``cuda_cases`` on the real kernels."""
    return [(n, "numpy") for n in numpy_cases] + [pytest.param(n, "cuda", marks=pytest.mark.gpu) for n in cuda_cases]


@pytest.mark.parametrize("name,kind", _variants(CASES))
def test_node_parallel_sweepers_standalone(tmp_path, name, kind):
    _check(tmp_path, name, kind, "own")


@pytest.mark.skipif(REF_PATHS is None, reason="reference tree not present (no /root/reference, no oracle/_ref)")
@pytest.mark.parametrize("name,kind", _variants(CASES[2:], CASES[2:]))
def test_node_parallel_sweepers_under_the_reference_controller(tmp_path, name, kind):
    _check(tmp_path, name, kind, "plugin")


@pytest.mark.skipif(REF_PATHS is None, reason="reference tree not present (no /root/reference, no oracle/_ref)")
@pytest.mark.parametrize("name,kind", _variants(CASES[1:2], CASES[:1]))
def test_reference_node_parallel_sweepers_on_the_facade(tmp_path, name, kind):
    _check(tmp_path, name, kind, "reference")


# ---------------------------------------------------------------------------------------------------------------------
# the reference's own test of its node-parallel sweepers (pySDC/tests/test_sweepers/test_MPI_sweeper.py:4-153): one
# controller step with the MPI sweeper and with the serial one must give the same end point and residual (1e-14), for
# node counts, node types, residual types, initial guesses, IMEX and a two-level (MLSDC) run that goes through the
# reference's unmodified base_transfer_MPI (transfer_classes/BaseTransferMPI.py) - here with the plug-in classes
# ---------------------------------------------------------------------------------------------------------------------
def _mpi_vs_serial_worker(rank, world, port, kind, cases, ref_paths, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SDCB200_CHECK_TAGS="1")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    for p in reversed(ref_paths):
        sys.path.insert(0, p)
    import pysdc_b200.mpi_facade

    sys.path.insert(0, pysdc_b200.mpi_facade.PATH)
    transport = "gloo"
    if kind == "cuda":
        import torch

        transport = "nccl" if torch.cuda.device_count() >= world else "gloo"
        torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group(transport, rank=rank, world_size=world)
    try:
        from mpi4py import MPI  # the facade
        from pySDC.implementations.controller_classes.controller_nonMPI import controller_nonMPI
        from pySDC.implementations.transfer_classes.BaseTransferMPI import base_transfer_MPI

        from pysdc_b200 import backend
        from pysdc_b200 import pysdc_plugin as plugin

        if kind == "cuda":
            backend.set_backend(backend.CudaBackend())
        else:
            from fake_backend import NumpyBackend

            backend.set_backend(NumpyBackend())

        def run(use_MPI, quad_type, residual_type, imex, init_guess, ML):
            name = "imex_1st_order" if imex else "generic_implicit"
            sp = dict(num_nodes=world, quad_type=quad_type, QI="IEpar", QE="PIC", initial_guess=init_guess)
            pp = dict(nvars=[(31, 31), (15, 15)] if ML > 1 else (31, 31), bc="dirichlet-zero", freq=(2, 2), nu=0.1,
                      solver_type="CG", lintol=1e-14, liniter=1000)
            d = dict(problem_class=plugin.heatNd_forced if imex else plugin.heatNd_unforced, problem_params=pp,
                     sweeper_class=getattr(plugin, name + ("_MPI" if use_MPI else "")), sweeper_params=sp,
                     level_params=dict(dt=1e-1, residual_type=residual_type), step_params=dict(maxiter=1))
            if use_MPI:
                sp["comm"] = MPI.COMM_WORLD
            if ML > 1:
                d.update(space_transfer_class=plugin.mesh_to_mesh)
                if use_MPI:
                    d["base_transfer_class"] = base_transfer_MPI
            c = controller_nonMPI(1, {"logger_level": 40}, d)
            P = c.MS[0].levels[0].prob
            u0 = P.u_exact(0) if imex else P.u_exact(0) + 1.0
            c.run(u0, 0, 1e-1)
            L = c.MS[0].levels[0]
            L.sweep.compute_end_point()
            return L.uend.get(), L.status.residual

        worst = 0.0
        for case in cases:
            u_mpi, r_mpi = run(True, *case)
            u_ser, r_ser = run(False, *case)
            assert np.allclose(u_mpi, u_ser, atol=1e-13), (case, float(np.max(np.abs(u_mpi - u_ser))))
            assert np.allclose(r_mpi, r_ser, atol=1e-13), (case, r_mpi, r_ser)
            worst = max(worst, float(np.max(np.abs(u_mpi - u_ser))))
        with open(os.path.join(out_dir, f"ok_{rank}"), "w") as f:
            f.write(repr(worst))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(REF_PATHS is None, reason="reference tree not present (no /root/reference, no oracle/_ref)")
@pytest.mark.parametrize("kind,world", [("numpy", 3), pytest.param("cuda", 2, marks=pytest.mark.gpu)])
def test_mpi_sweeper_equals_serial_sweeper_like_the_reference_test(tmp_path, kind, world):
    cases = []  # (quad_type, residual_type, imex, initial_guess, ML) as in test_MPI_sweeper.py:156-205
    for quad_type in ("GAUSS", "RADAU-RIGHT"):
        for residual_type in ("last_abs", "full_rel"):
            for imex in (False, True):
                cases.append((quad_type, residual_type, imex, "spread", 1))
    cases += [("RADAU-RIGHT", "full_abs", False, "copy", 1), ("RADAU-RIGHT", "full_abs", True, "zero", 1),
              ("RADAU-RIGHT", "full_abs", False, "spread", 2), ("RADAU-RIGHT", "last_rel", True, "spread", 2)]
    if kind == "cuda":
        cases = cases[::3] + cases[-2:]
    mp.spawn(_mpi_vs_serial_worker, args=(world, free_port(), kind, cases, REF_PATHS, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(tmp_path, f"ok_{r}")) for r in range(world))


# ---------------------------------------------------------------------------------------------------------------------
# ... and the reference's test ITSELF: pySDC/tests/test_sweepers/test_MPI_sweeper.py::individual_test(launch=False, ...)
# called, unmodified, in every rank for the reference's own parameter grid (2 nodes x {GAUSS, RADAU-RIGHT} x {last_abs,
# full_rel} x {imex, not} x {spread, copy, zero} x ML in {1, 2, 3}: 72 combinations; ML > 1 goes through the reference's
# base_transfer_MPI), with the class names resolving to the plug-in classes (see tests/test_reference_suite.py)
# ---------------------------------------------------------------------------------------------------------------------
def _reference_mpi_test_worker(rank, world, port, kind, ref_paths, out_dir, stride, which="sweeper"):
    import importlib
    import itertools
    import types

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    for p in reversed(ref_paths):
        sys.path.insert(0, p)
    import pysdc_b200.mpi_facade

    sys.path.insert(0, pysdc_b200.mpi_facade.PATH)
    transport = "gloo"
    if kind == "cuda":
        import torch

        transport = "nccl" if torch.cuda.device_count() >= world else "gloo"
        torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group(transport, rank=rank, world_size=world)
    try:
        from pysdc_b200 import backend
        from pysdc_b200 import pysdc_plugin as plugin

        if kind == "cuda":
            backend.set_backend(backend.CudaBackend())
        else:
            from fake_backend import NumpyBackend

            backend.set_backend(NumpyBackend())
        base = "pySDC.implementations."
        for name, exports in {
                base + "problem_classes.HeatEquation_ND_FD": ["heatNd_unforced", "heatNd_forced"],
                base + "sweeper_classes.generic_implicit": ["generic_implicit"],
                base + "sweeper_classes.imex_1st_order": ["imex_1st_order"],
                base + "sweeper_classes.generic_implicit_MPI": ["generic_implicit_MPI"],
                base + "sweeper_classes.imex_1st_order_MPI": ["imex_1st_order_MPI"],
                base + "transfer_classes.TransferMesh": ["mesh_to_mesh"]}.items():
            mod = types.ModuleType(name)
            for e in exports:
                setattr(mod, e, getattr(plugin, e))
            sys.modules[name] = mod
        if which == "base_transfer":
            # pySDC/tests/test_transfer_classes/test_base_transfer_MPI.py:69-117: restrict / prolong / prolong_f of the
            # reference's base_transfer_MPI (node-parallel levels) against its serial BaseTransfer, field by field
            ref_test = importlib.import_module("pySDC.tests.test_transfer_classes.test_base_transfer_MPI")
            launches0 = backend.get_backend().launches
            for nvars in (32, 16):  # :46
                ref_test._test_MPI_nonMPI_consistency(nvars)
            assert backend.get_backend().launches > launches0
            with open(os.path.join(out_dir, f"ok_{rank}"), "w") as f:
                f.write("2")
            return
        ref_test = importlib.import_module("pySDC.tests.test_sweepers.test_MPI_sweeper")
        grid = list(itertools.product(["GAUSS", "RADAU-RIGHT"], ["last_abs", "full_rel"], [True, False],
                                      ["spread", "copy", "zero"], [1, 2, 3]))
        launches0 = backend.get_backend().launches
        if kind == "cuda":  # the real kernels run the 512-point single-level cases; the 2- and 4-point grids of the
            grid = [c for c in grid if c[4] == 1]  # reference's multi-level cases are covered on the numpy double
        for quad_type, residual_type, imex, init_guess, ML in grid[::stride]:
            ref_test.individual_test(launch=False, num_nodes=world, quad_type=quad_type, residual_type=residual_type,
                                     imex=imex, init_guess=init_guess, useNCCL=False, ML=ML)
        assert backend.get_backend().launches > launches0
        with open(os.path.join(out_dir, f"ok_{rank}"), "w") as f:
            f.write(str(len(grid[::stride])))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(REF_PATHS is None, reason="reference tree not present (no /root/reference, no oracle/_ref)")
@pytest.mark.parametrize("kind,stride", [("numpy", 1), pytest.param("cuda", 3, marks=pytest.mark.gpu)])
def test_reference_test_MPI_sweeper_passes_on_plugin_classes(tmp_path, kind, stride):
    world = 2  # test_MPI_sweeper.py:141 (num_nodes = 2)
    mp.spawn(_reference_mpi_test_worker, args=(world, free_port(), kind, REF_PATHS, str(tmp_path), stride), nprocs=world,
             join=True)
    assert all(os.path.exists(os.path.join(tmp_path, f"ok_{r}")) for r in range(world))


@pytest.mark.skipif(REF_PATHS is None, reason="reference tree not present (no /root/reference, no oracle/_ref)")
@pytest.mark.parametrize("kind,world", [("numpy", 2), ("numpy", 3), pytest.param("cuda", 2, marks=pytest.mark.gpu)])
def test_reference_test_base_transfer_MPI_passes_on_plugin_classes(tmp_path, kind, world):
    mp.spawn(_reference_mpi_test_worker, args=(world, free_port(), kind, REF_PATHS, str(tmp_path), 1, "base_transfer"),
             nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(tmp_path, f"ok_{r}")) for r in range(world))
