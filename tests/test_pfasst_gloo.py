"""Time-parallel controller (pysdc_b200.pfasst.controller_MPI) on CPU: one process per time slice over `gloo`, kernel
library replaced by the numpy test double.  Expected values are the fixtures the UNMODIFIED reference produced with its
virtual-parallel controller (oracle/make_golden.py: pfasst_*), including the reference's own known answer
tutorial/step_8/A_visualize_residuals.py:56-58 (7 iterations on each of the 8 slices).  Also: transfer operators vs the
formulas of helpers/transfer_helper.py, and serial MLSDC through the same controller."""
import json
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, free_port, load_golden


def _description(name):
    from pysdc_b200 import problems, sweepers
    from pysdc_b200.transfer import mesh_to_mesh

    if name == "pfasst_step8A_heat1d":
        d = dict(problem_class=problems.heatNd_unforced,
                 problem_params=dict(nu=0.1, freq=2, nvars=[63, 31], bc="dirichlet-zero"),
                 sweeper_class=sweepers.generic_implicit,
                 sweeper_params=dict(quad_type="RADAU-RIGHT", num_nodes=[3], QI="LU"),
                 level_params=dict(restol=5e-10, dt=0.125), step_params=dict(maxiter=50, errtol=1e-5),
                 space_transfer_class=mesh_to_mesh, space_transfer_params=dict(rorder=2, iorder=6))
        return d, dict(logger_level=40, all_to_done=True, predict_type="pfasst_burnin"), 0.0, 1.0, 8
    spec, _ = load_golden(name)
    pp = dict(spec["problem_params"])
    pp["nvars"] = [tuple(v) for v in pp["nvars"]]
    pp["freq"] = tuple(pp["freq"])
    d = dict(problem_class=getattr(problems, spec["problem"]), problem_params=pp,
             sweeper_class=getattr(sweepers, spec["sweeper"]), sweeper_params=dict(spec["sweeper_params"]),
             level_params=dict(spec["level_params"]), step_params=dict(spec["step_params"]),
             space_transfer_class=mesh_to_mesh, space_transfer_params=dict(spec["space_transfer_params"]))
    return d, dict(spec["controller_params"]), spec["t0"], spec["Tend"], spec["num_procs"]


def _worker(rank, world, port, name, out_dir, Tend_override=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SDCB200_CHECK_TAGS="1")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fake_backend import NumpyBackend
        from pysdc_b200 import backend
        from pysdc_b200.parallel import TorchComm
        from pysdc_b200.pfasst import controller_MPI
        from pysdc_b200.stats import get_sorted

        backend.set_backend(NumpyBackend())
        d, cp, t0, Tend, _ = _description(name)
        Tend = Tend if Tend_override is None else Tend_override
        c = controller_MPI(cp, d, comm=TorchComm())
        P = c.S.levels[0].prob
        uend, stats = c.run(u0=P.u_exact(t0), t0=t0, Tend=Tend)
        niter = [(float(t), int(v)) for t, v in get_sorted(stats, type="niter", sortby="time")]
        with open(os.path.join(out_dir, f"niter_{rank}.json"), "w") as f:
            json.dump(niter, f)
        np.save(os.path.join(out_dir, f"uend_{rank}.npy"), uend.get())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,world", [("pfasst_heat2d_imex_63_p4", 4), ("pfasst_step8A_heat1d", 8),
                                        ("pfasst_heat2d_imex_63_p4", 2)])
def test_pfasst_matches_reference_fixture(tmp_path, name, world):
    """world == num_procs of the fixture: one block, iteration counts and end value must match the reference's virtual
    PFASST run.  world == 2 < num_procs: two blocks of two slices - a different (shorter-pipeline) PFASST schedule, so
    only the end value is compared (to the discretisation-independent tolerance of the fixture's restol)."""
    _, g = load_golden(name)
    mp.spawn(_worker, args=(world, free_port(), name, str(tmp_path)), nprocs=world, join=True)
    niter = []
    for r in range(world):
        niter += [tuple(x) for x in json.load(open(os.path.join(tmp_path, f"niter_{r}.json")))]
    niter = [v for _, v in sorted(niter)]
    uend = np.load(os.path.join(tmp_path, "uend_0.npy"))
    for r in range(1, world):
        assert np.array_equal(np.load(os.path.join(tmp_path, f"uend_{r}.npy")), uend)  # block-end broadcast
    scale = np.max(np.abs(g["uend"]))
    if world == len(g["niter"]):
        assert niter == g["niter"].tolist()
        assert np.max(np.abs(uend - g["uend"])) / scale < 1e-10
    else:
        assert len(niter) == len(g["niter"])
        assert np.max(np.abs(uend - g["uend"])) / scale < 1e-7


def test_transfer_operators():
    """Prolongation reproduces polynomials up to its order away from and at the Dirichlet boundary; restriction is the
    scaled transpose (TransferMesh.py:72-90); the N-D device application equals the Kronecker product."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_backend import NumpyBackend
    from pysdc_b200 import backend, problems
    from pysdc_b200.transfer import interpolation_matrix_1d, mesh_to_mesh

    old = backend._backend
    backend.set_backend(NumpyBackend())
    try:
        nf, nc = 31, 15
        fg = np.array([(i + 1) / (nf + 1) for i in range(nf)])
        cg = np.array([(i + 1) / (nc + 1) for i in range(nc)])
        for k in (2, 4, 6):
            P1 = interpolation_matrix_1d(fg, cg, k=k)
            assert P1.shape == (nf, nc)
            # odd fine points coincide with coarse points: injection
            assert np.array_equal(P1[1::2], np.eye(nc))
            # polynomials vanishing at both walls of degree < k are reproduced exactly
            for deg in range(1, k - 1):
                poly = lambda x: x * (1 - x) * x ** (deg - 1)  # noqa: E731
                assert np.max(np.abs(P1 @ poly(cg) - poly(fg))) < 1e-13
        Pf = problems.heatNd_forced(nvars=(nf, nf), nu=0.1, freq=(2, 2), bc="dirichlet-zero", solver_type="CG")
        Pc = problems.heatNd_forced(nvars=(nc, nc), nu=0.1, freq=(2, 2), bc="dirichlet-zero", solver_type="CG")
        T = mesh_to_mesh(Pf, Pc, dict(rorder=2, iorder=6))
        rng = np.random.default_rng(3)
        G = Pc.dtype_f(Pc.init)
        g0, g1 = rng.standard_normal((nc, nc)), rng.standard_normal((nc, nc))
        G.impl[:] = g0
        G.expl[:] = g1
        F = T.prolong(G)
        assert type(F) is Pf.dtype_f
        K = np.kron(T.Pspace_1d, T.Pspace_1d)
        assert np.max(np.abs(F.impl.get() - (K @ g0.ravel()).reshape(nf, nf))) < 1e-13
        assert np.max(np.abs(F.expl.get() - (K @ g1.ravel()).reshape(nf, nf))) < 1e-13
        u = Pf.dtype_u(Pf.init)
        f0 = rng.standard_normal((nf, nf))
        u[:] = f0
        R = np.kron(T.Rspace_1d, T.Rspace_1d)
        assert np.max(np.abs(T.restrict(u).get() - (R @ f0.ravel()).reshape(nc, nc))) < 1e-13
        assert np.allclose(T.Rspace_1d[3, 5:10], [0.0, 0.25, 0.5, 0.25, 0.0])
    finally:
        backend.set_backend(old)


def test_partial_last_block_and_serial_mlsdc(tmp_path):
    """6 time steps on 4 ranks: one full block and a last block of two slices (sub-communicator of the first two
    ranks), compared with serial MLSDC through the same controller on a one-process communicator."""
    name, world, Tend = "pfasst_heat2d_imex_63_p4", 4, 1.5
    mp.spawn(_worker, args=(world, free_port(), name, str(tmp_path), Tend), nprocs=world, join=True)
    niter = []
    for r in range(world):
        niter += [tuple(x) for x in json.load(open(os.path.join(tmp_path, f"niter_{r}.json")))]
    assert len(niter) == 6 and sorted(t for t, _ in niter) == pytest.approx([0.0, 0.25, 0.5, 0.75, 1.0, 1.25])
    uend = np.load(os.path.join(tmp_path, "uend_0.npy"))  # ranks 0 and 1 hold the end value of the last block
    assert np.array_equal(np.load(os.path.join(tmp_path, "uend_1.npy")), uend)

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fake_backend import NumpyBackend
    from pysdc_b200 import backend
    from pysdc_b200.pfasst import controller_MPI

    old = backend._backend
    backend.set_backend(NumpyBackend())
    try:
        d, cp, t0, _, _ = _description(name)
        c = controller_MPI(cp, d)  # LocalComm: serial MLSDC
        P = c.S.levels[0].prob
        ref, stats = c.run(u0=P.u_exact(t0), t0=t0, Tend=Tend)
        assert np.max(np.abs(uend - ref.get())) / np.max(np.abs(ref.get())) < 1e-7
        assert abs(P.u_exact(Tend) - ref) < 5e-3  # second-order FD error of the 63^2 grid
    finally:
        backend.set_backend(old)
