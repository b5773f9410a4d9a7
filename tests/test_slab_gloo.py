"""World-size-2 (and 3) `gloo` tests of the multi-GPU HOST logic on CPU: slab partition, slab datatypes, halo-plane
exchange, global reductions, and a full SDC step through the slab-decomposed problem / sweeper classes compared with the
single-process run.  The kernel library is replaced by the numpy test double (tests/fake_backend.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _spec(n):
    return dict(problem_params=dict(nvars=(n, n, n), nu=0.1, freq=(1, 1, 1), bc="dirichlet-zero", solver_type="CG",
                                    lintol=1e-12, liniter=10000),
                sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="MIN-SR-NS", initial_guess="spread"),
                level_params=dict(dt=1e-3, restol=1e-9), step_params=dict(maxiter=20))


def _run_step(n, comm, u0_global):
    from pysdc_b200.controller import controller_nonMPI
    from pysdc_b200.problems import heatNd_unforced
    from pysdc_b200.stats import get_sorted
    from pysdc_b200.sweepers import generic_implicit

    sp = _spec(n)
    pp = dict(sp["problem_params"])
    if comm is not None:
        pp["comm"] = comm
    c = controller_nonMPI(1, {"logger_level": 40}, dict(
        problem_class=heatNd_unforced, problem_params=pp, sweeper_class=generic_implicit,
        sweeper_params=sp["sweeper_params"], level_params=sp["level_params"], step_params=sp["step_params"]))
    P = c.MS[0].levels[0].prob
    u0 = P.dtype_u(P.init)
    u0[:] = u0_global
    uend, stats = c.run(u0=u0, t0=0.0, Tend=1e-3)
    niter = [v for _, v in get_sorted(stats, type="niter")]
    return uend, niter, P


def _run_imex(n, comm):
    """Forced heat, IMEX sweeper with LU (sequential node solves) - the multi-component datatype on slabs."""
    from pysdc_b200.controller import controller_nonMPI
    from pysdc_b200.problems import heatNd_forced
    from pysdc_b200.stats import get_sorted
    from pysdc_b200.sweepers import imex_1st_order

    pp = dict(_spec(n)["problem_params"], freq=(2, 2, 2))
    if comm is not None:
        pp["comm"] = comm
    c = controller_nonMPI(1, {"logger_level": 40}, dict(
        problem_class=heatNd_forced, problem_params=pp, sweeper_class=imex_1st_order,
        sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="LU"),
        level_params=dict(dt=5e-3, restol=1e-9), step_params=dict(maxiter=50)))
    P = c.MS[0].levels[0].prob
    uend, stats = c.run(u0=P.u_exact(0.0), t0=0.0, Tend=1e-2)
    return uend, [v for _, v in get_sorted(stats, type="niter")], P


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_backend import NumpyBackend
        from pysdc_b200 import backend
        from pysdc_b200.datatypes import mesh
        from pysdc_b200.parallel import SlabComm, split_planes

        backend.set_backend(NumpyBackend())
        comm = SlabComm()
        counts = split_planes(n, world)
        lay = comm.slab_layout((n, n, n))
        assert lay.nz == counts[rank] and lay.z0 == sum(counts[:rank])

        # datatype: global assignment is cut to the slab, abs() is global, gather() reassembles
        rng = np.random.default_rng(7)
        g = rng.standard_normal((n, n, n))
        m = mesh(((n, n, n), comm, np.dtype("float64")))
        m[:] = g
        assert m.shape == (lay.nz, n, n)
        assert np.array_equal(m.get(), g[lay.z0: lay.z0 + lay.nz])
        assert abs(m) == np.max(np.abs(g))
        assert np.array_equal(m.gather(), g)
        assert abs(m - m) == 0.0

        # halo exchange: neighbours' boundary planes arrive, domain ends keep zeros
        comm.exchange_halos([m])
        full = NumpyBackend._slab_full(lay, m.flat).numpy()
        lo = g[lay.z0 - 1] if rank > 0 else np.zeros((n, n))
        hi = g[lay.z0 + lay.nz] if rank + 1 < world else np.zeros((n, n))
        assert np.array_equal(full[0], lo) and np.array_equal(full[-1], hi)

        # preconditioned distributed solve == plain distributed solve (to the solver tolerance), fewer iterations
        from pysdc_b200.problems import heatNd_unforced

        kw = dict(_spec(n)["problem_params"], comm=comm)
        Pa, Pb = heatNd_unforced(**kw), heatNd_unforced(**kw, preconditioner="chebyshev")
        rhs_g = rng.standard_normal((n, n, n))

        def solve(Pr):
            b, x = Pr.dtype_u(Pr.init), Pr.dtype_u(Pr.init)
            b[:] = rhs_g
            x[:] = g
            Pr.solve_system_batch([b], [2e-3], [x])
            return x.gather(), Pr.work_counters["CG"].niter

        (xa, ia), (xb_, ib) = solve(Pa), solve(Pb)
        assert np.max(np.abs(xa - xb_)) / np.max(np.abs(xa)) < 1e-10 and ib < ia

        # a full SDC step on slabs == the single-process step
        u0 = np.random.default_rng(1234).standard_normal((n, n, n))
        uend, niter, P = _run_step(n, comm, u0)
        np.save(os.path.join(out_dir, f"uend_{rank}.npy"), uend.gather())
        np.save(os.path.join(out_dir, f"niter_{rank}.npy"), np.array(niter))
        np.save(os.path.join(out_dir, f"cg_{rank}.npy"), np.array(P.work_counters["CG"].niter))
        # output of slab fields (fields_io.py): every rank writes its planes into the global record of a FieldsIO file
        path = os.path.join(out_dir, "slab.pySDC")
        out = P.getOutputFile(path)
        out.addField(0.0, P.processSolutionForOutput(P.u_exact(0.0)))
        out.addField(1e-3, P.processSolutionForOutput(uend))
        out.flush()
        uend, niter, P = _run_imex(n, comm)
        assert abs(P.u_exact(1e-2) - uend) < 0.05  # global max-norm through the communicator
        np.save(os.path.join(out_dir, f"imex_uend_{rank}.npy"), uend.gather())
        np.save(os.path.join(out_dir, f"imex_niter_{rank}.npy"), np.array(niter))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 15), (3, 13)])
def test_slab_step_matches_serial(tmp_path, world, n):
    from conftest import free_port

    port = free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_backend import NumpyBackend
    from pysdc_b200 import backend
    from pysdc_b200.datatypes import mesh

    old, old_comm = backend._backend, mesh.comm
    backend.set_backend(NumpyBackend())
    try:
        mesh.comm = None
        u0 = np.random.default_rng(1234).standard_normal((n, n, n))
        uend, niter, P = _run_step(n, None, u0)
        ref = uend.get()
        for r in range(world):
            got = np.load(os.path.join(tmp_path, f"uend_{r}.npy"))
            assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-12
            assert list(np.load(os.path.join(tmp_path, f"niter_{r}.npy"))) == niter
            assert abs(int(np.load(os.path.join(tmp_path, f"cg_{r}.npy"))) - P.work_counters["CG"].niter) <= 2
        # the file the ranks wrote together holds the GLOBAL fields (readable without any communicator)
        from pysdc_b200.fields_io import RectilinearFile

        f = RectilinearFile.fromFile(os.path.join(tmp_path, "slab.pySDC"))
        assert f.times == [0.0, 1e-3] and f.gridSizes == [n, n, n]
        assert np.array_equal(f.readField(0)[1][0], P.u_exact(0.0).get())
        assert np.max(np.abs(f.readField(1)[1][0] - ref)) / np.max(np.abs(ref)) < 1e-12
        uend, niter, _ = _run_imex(n, None)
        ref = uend.get()
        for r in range(world):
            got = np.load(os.path.join(tmp_path, f"imex_uend_{r}.npy"))
            assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-12
            assert list(np.load(os.path.join(tmp_path, f"imex_niter_{r}.npy"))) == niter
    finally:
        backend.set_backend(old)
        mesh.comm = old_comm


def test_split_planes():
    from pysdc_b200.parallel import split_planes

    assert split_planes(511, 8) == [64] * 7 + [63]
    assert split_planes(511, 1) == [511]
    assert sum(split_planes(127, 3)) == 127


# ---------------------------------------------------------------------------------------------------------------------
# nodes x slabs: the node-parallel sweeper (one collocation node per outer rank) on slab-decomposed fields, i.e. a
# 2-D process grid built with parallel.cartesian_comms - against the single-process run
# ---------------------------------------------------------------------------------------------------------------------
def _grid_worker(rank, world, port, n, n_nodes, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SDCB200_CHECK_TAGS="1")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_backend import NumpyBackend
        from pysdc_b200 import backend
        from pysdc_b200.controller import controller_nonMPI
        from pysdc_b200.parallel import cartesian_comms
        from pysdc_b200.problems import heatNd_unforced
        from pysdc_b200.stats import get_sorted
        from pysdc_b200.sweepers import generic_implicit_MPI

        backend.set_backend(NumpyBackend())
        node_comm, space_comm = cartesian_comms(n_nodes, world // n_nodes)
        assert node_comm.size == n_nodes and space_comm.size == world // n_nodes
        assert node_comm.rank == rank // space_comm.size and space_comm.rank == rank % space_comm.size
        sp = _spec(n)
        c = controller_nonMPI(1, {"logger_level": 40}, dict(
            problem_class=heatNd_unforced, problem_params=dict(sp["problem_params"], comm=space_comm),
            sweeper_class=generic_implicit_MPI,
            sweeper_params=dict(sp["sweeper_params"], num_nodes=n_nodes, comm=node_comm),
            level_params=sp["level_params"], step_params=sp["step_params"]))
        P = c.MS[0].levels[0].prob
        u0 = P.dtype_u(P.init)
        u0[:] = np.random.default_rng(1234).standard_normal((n, n, n))
        uend, stats = c.run(u0=u0, t0=0.0, Tend=2e-3)
        np.save(os.path.join(out_dir, f"uend_{rank}.npy"), uend.gather())
        np.save(os.path.join(out_dir, f"niter_{rank}.npy"), np.array([v for _, v in get_sorted(stats, type="niter")]))
    finally:
        dist.destroy_process_group()


def test_nodes_times_slabs_matches_serial(tmp_path):
    from conftest import free_port

    n, n_nodes, world = 13, 2, 4
    mp.spawn(_grid_worker, args=(world, free_port(), n, n_nodes, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_backend import NumpyBackend
    from pysdc_b200 import backend
    from pysdc_b200.controller import controller_nonMPI
    from pysdc_b200.datatypes import mesh
    from pysdc_b200.problems import heatNd_unforced
    from pysdc_b200.stats import get_sorted
    from pysdc_b200.sweepers import generic_implicit

    old, old_comm = backend._backend, mesh.comm
    backend.set_backend(NumpyBackend())
    try:
        mesh.comm = None
        sp = _spec(n)
        c = controller_nonMPI(1, {"logger_level": 40}, dict(
            problem_class=heatNd_unforced, problem_params=sp["problem_params"], sweeper_class=generic_implicit,
            sweeper_params=dict(sp["sweeper_params"], num_nodes=n_nodes), level_params=sp["level_params"],
            step_params=sp["step_params"]))
        P = c.MS[0].levels[0].prob
        u0 = P.dtype_u(P.init)
        u0[:] = np.random.default_rng(1234).standard_normal((n, n, n))
        uend, stats = c.run(u0=u0, t0=0.0, Tend=2e-3)
        ref, niter = uend.get(), [v for _, v in get_sorted(stats, type="niter")]
        for r in range(world):
            got = np.load(os.path.join(tmp_path, f"uend_{r}.npy"))
            assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-12
            assert list(np.load(os.path.join(tmp_path, f"niter_{r}.npy"))) == niter
    finally:
        backend.set_backend(old)
        mesh.comm = old_comm


# ---------------------------------------------------------------------------------------------------------------------
# time slices x slabs: the time-parallel controller (one time slice per outer rank, uend -> u[0] hand-over between the
# slices rank by rank of the space communicator) on slab-decomposed fields
# ---------------------------------------------------------------------------------------------------------------------
def _time_space_worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SDCB200_CHECK_TAGS="1")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_backend import NumpyBackend
        from pysdc_b200 import backend
        from pysdc_b200.parallel import cartesian_comms
        from pysdc_b200.pfasst import controller_MPI
        from pysdc_b200.problems import heatNd_unforced
        from pysdc_b200.stats import get_sorted
        from pysdc_b200.sweepers import generic_implicit

        backend.set_backend(NumpyBackend())
        time_comm, space_comm = cartesian_comms(2, world // 2)
        sp = _spec(n)
        c = controller_MPI({"logger_level": 40}, dict(
            problem_class=heatNd_unforced, problem_params=dict(sp["problem_params"], comm=space_comm),
            sweeper_class=generic_implicit, sweeper_params=dict(sp["sweeper_params"], QI="LU"),
            level_params=sp["level_params"], step_params=sp["step_params"]), comm=time_comm)
        P = c.S.levels[0].prob
        u0 = P.dtype_u(P.init)
        u0[:] = np.random.default_rng(1234).standard_normal((n, n, n))
        uend, stats = c.run(u0=u0, t0=0.0, Tend=4e-3)  # two blocks of two slices
        np.save(os.path.join(out_dir, f"uend_{rank}.npy"), uend.gather())
        np.save(os.path.join(out_dir, f"times_{rank}.npy"), np.array([t for t, _ in get_sorted(stats, type="niter", sortby="time")]))
    finally:
        dist.destroy_process_group()


def test_time_slices_times_slabs_matches_serial_time_stepping(tmp_path):
    from conftest import free_port

    n, world = 13, 4
    mp.spawn(_time_space_worker, args=(world, free_port(), n, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_backend import NumpyBackend
    from pysdc_b200 import backend
    from pysdc_b200.controller import controller_nonMPI
    from pysdc_b200.datatypes import mesh
    from pysdc_b200.problems import heatNd_unforced
    from pysdc_b200.sweepers import generic_implicit

    old, old_comm = backend._backend, mesh.comm
    backend.set_backend(NumpyBackend())
    try:
        mesh.comm = None
        sp = _spec(n)
        c = controller_nonMPI(1, {"logger_level": 40}, dict(
            problem_class=heatNd_unforced, problem_params=sp["problem_params"], sweeper_class=generic_implicit,
            sweeper_params=dict(sp["sweeper_params"], QI="LU"), level_params=sp["level_params"],
            step_params=sp["step_params"]))
        P = c.MS[0].levels[0].prob
        u0 = P.dtype_u(P.init)
        u0[:] = np.random.default_rng(1234).standard_normal((n, n, n))
        ref = c.run(u0=u0, t0=0.0, Tend=4e-3)[0].get()
        got = [np.load(os.path.join(tmp_path, f"uend_{r}.npy")) for r in range(world)]
        for r in range(world):
            assert np.array_equal(got[r], got[0])  # block-end broadcast: the same end value on every rank
            # time-parallel SDC iterates every step to restol = 1e-9: the end values agree to that level, not to rounding
            assert np.max(np.abs(got[r] - ref)) / np.max(np.abs(ref)) < 1e-8
        # slice 0 (ranks 0, 1) took steps 0 and 2, slice 1 (ranks 2, 3) steps 1 and 3
        assert np.allclose(np.load(os.path.join(tmp_path, "times_0.npy")), [0.0, 2e-3])
        assert np.allclose(np.load(os.path.join(tmp_path, "times_3.npy")), [1e-3, 3e-3])
    finally:
        backend.set_backend(old)
        mesh.comm = old_comm
