"""The reference's OWN time-parallel controller (controller_MPI, unmodified) on the torch.distributed-backed mpi4py
facade, driving the plug-in classes.  ``numpy`` variant: gloo + the numpy test double (CPU suite); ``cuda`` variant
(``-m gpu``): the real CUDA kernels, one process per time slice (NCCL when the box has a GPU per slice, else the slices
share the device and hand over through gloo).  The reference comes from /root/reference or the shipped oracle/_ref."""
import json
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, free_port, load_golden, reference_paths

REF_PATHS = reference_paths()
pytestmark = pytest.mark.skipif(REF_PATHS is None, reason="reference tree not present (no /root/reference, no oracle/_ref)")


def _worker(rank, world, port, name, out_dir, kind, ref_paths):
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
        from pysdc_b200 import backend

        if kind == "cuda":
            backend.set_backend(backend.CudaBackend())
        else:
            from fake_backend import NumpyBackend

            backend.set_backend(NumpyBackend())
        from pySDC.helpers.stats_helper import get_sorted
        from pySDC.implementations.controller_classes.controller_MPI import controller_MPI

        from pysdc_b200 import pysdc_plugin as plugin
        from pysdc_b200.transfer import mesh_to_mesh

        spec, _ = load_golden(name)
        pp = dict(spec["problem_params"])
        pp["nvars"] = [tuple(v) for v in pp["nvars"]]
        pp["freq"] = tuple(pp["freq"])
        d = dict(problem_class=getattr(plugin, spec["problem"]), problem_params=pp,
                 sweeper_class=getattr(plugin, spec["sweeper"]), sweeper_params=dict(spec["sweeper_params"]),
                 level_params=dict(spec["level_params"]), step_params=dict(spec["step_params"]),
                 space_transfer_class=mesh_to_mesh, space_transfer_params=dict(spec["space_transfer_params"]))
        c = controller_MPI(controller_params=dict(spec["controller_params"]), description=d, comm=MPI.COMM_WORLD)
        P = c.S.levels[0].prob
        uend, stats = c.run(u0=P.u_exact(0.0), t0=0.0, Tend=spec["Tend"])
        niter = [(float(t), int(v)) for t, v in get_sorted(stats, type="niter", sortby="time")]
        with open(os.path.join(out_dir, f"niter_{rank}.json"), "w") as f:
            json.dump(niter, f)
        np.save(os.path.join(out_dir, f"uend_{rank}.npy"), uend.get())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["numpy", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_reference_controller_MPI_on_the_facade(tmp_path, kind):
    name, world = "pfasst_heat2d_imex_63_p4", 4
    _, g = load_golden(name)
    mp.spawn(_worker, args=(world, free_port(), name, str(tmp_path), kind, REF_PATHS), nprocs=world, join=True)
    niter = []
    for r in range(world):
        niter += [tuple(x) for x in json.load(open(os.path.join(tmp_path, f"niter_{r}.json")))]
    assert [v for _, v in sorted(niter)] == g["niter"].tolist()
    uend = np.load(os.path.join(tmp_path, "uend_0.npy"))
    assert np.max(np.abs(uend - g["uend"])) / np.max(np.abs(g["uend"])) < 1e-10
