#!/usr/bin/env python
"""Fixtures of the reference's OWN node-parallel sweepers (``generic_implicit_MPI``, ``imex_1st_order_MPI``:
sweeper_classes/generic_implicit_MPI.py, imex_1st_order_MPI.py), unmodified, one process per collocation node.

TEST INFRASTRUCTURE, build container only (needs /root/reference).  mpi4py is absent from the image, so the reference's
``from mpi4py import MPI`` resolves to ``pysdc_b200/mpi_facade`` — a transport shim (rank / size, Reduce / Allreduce /
Bcast / allreduce on torch.distributed's gloo backend); sweeper, problem, datatype, controller and collocation code are
the reference's.  Every case is also run with the reference's SERIAL sweeper in the same process group: the two must
give the same iteration counts (asserted before anything is written), which pins the fixtures on two independent
code paths of the reference.

    python oracle/make_golden_node_parallel.py
"""
import json
import os
import socket
import sys
import tempfile

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("PYSDC_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # the serial fixture run_heat3d_gi_minsrns_31 (oracle/make_golden.py), one node per rank
    "nodepar_heat3d_gi_minsrns_31": dict(
        problem="heatNd_unforced", sweeper="generic_implicit",
        problem_params=dict(nvars=[31, 31, 31], nu=0.1, freq=[1, 1, 1], bc="dirichlet-zero", solver_type="CG",
                            lintol=1e-12, liniter=10000),
        sweeper_params=dict(num_nodes=4, quad_type="RADAU-RIGHT", QI="MIN-SR-NS", initial_guess="spread"),
        level_params=dict(dt=1e-3, restol=1e-8), step_params=dict(maxiter=50), t0=0.0, Tend=2e-3, u0="random", seed=1234),
    # quadrature end-point update (Allreduce path, generic_implicit_MPI.py:256-266) with a relative residual
    "nodepar_heat2d_gi_minsrs_collupdate_63": dict(
        problem="heatNd_unforced", sweeper="generic_implicit",
        problem_params=dict(nvars=[63, 63], nu=0.1, freq=[2, 2], bc="dirichlet-zero", solver_type="CG", lintol=1e-12,
                            liniter=10000),
        sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="MIN-SR-S", initial_guess="copy",
                            do_coll_update=True),
        level_params=dict(dt=0.01, restol=1e-9, residual_type="full_rel"), step_params=dict(maxiter=50), t0=0.0,
        Tend=0.03, u0="exact"),
    # IMEX, Picard explicit part (the only one imex_1st_order_MPI supports, :9-12)
    "nodepar_heat2d_imex_minsrs_pic_63": dict(
        problem="heatNd_forced", sweeper="imex_1st_order",
        problem_params=dict(nvars=[63, 63], nu=0.1, freq=[4, 4], bc="dirichlet-zero", solver_type="CG", lintol=1e-12,
                            liniter=10000),
        sweeper_params=dict(num_nodes=3, quad_type="RADAU-RIGHT", QI="MIN-SR-S", QE="PIC"),
        level_params=dict(dt=0.02, restol=1e-10, residual_type="last_abs"), step_params=dict(maxiter=50), t0=0.0,
        Tend=0.04, u0="exact"),
}


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(spec, parallel, comm):
    from pySDC.helpers.stats_helper import get_sorted
    from pySDC.implementations.controller_classes.controller_nonMPI import controller_nonMPI
    from pySDC.implementations.hooks.log_work import LogWork
    from pySDC.implementations.problem_classes.HeatEquation_ND_FD import heatNd_forced, heatNd_unforced
    from pySDC.implementations.sweeper_classes.generic_implicit import generic_implicit
    from pySDC.implementations.sweeper_classes.generic_implicit_MPI import generic_implicit_MPI
    from pySDC.implementations.sweeper_classes.imex_1st_order import imex_1st_order
    from pySDC.implementations.sweeper_classes.imex_1st_order_MPI import imex_1st_order_MPI

    probs = dict(heatNd_unforced=heatNd_unforced, heatNd_forced=heatNd_forced)
    sweeps = dict(generic_implicit=(generic_implicit, generic_implicit_MPI), imex_1st_order=(imex_1st_order, imex_1st_order_MPI))
    pp = dict(spec["problem_params"])
    for k in ("nvars", "freq"):
        pp[k] = tuple(pp[k])
    sp = dict(spec["sweeper_params"])
    if parallel:
        sp["comm"] = comm
    d = dict(problem_class=probs[spec["problem"]], problem_params=pp, sweeper_class=sweeps[spec["sweeper"]][int(parallel)],
             sweeper_params=sp, level_params=dict(spec["level_params"]), step_params=dict(spec["step_params"]))
    c = controller_nonMPI(num_procs=1, controller_params={"logger_level": 40, "hook_class": [LogWork]}, description=d)
    P = c.MS[0].levels[0].prob
    if spec["u0"] == "exact":
        u0 = P.u_exact(spec["t0"])
    else:
        u0 = P.u_init
        u0[:] = np.random.default_rng(spec["seed"]).standard_normal(u0.shape)
    uend, stats = c.run(u0=u0, t0=spec["t0"], Tend=spec["Tend"])
    its = get_sorted(stats, type="niter", sortby="time")
    res = [[float(v) for _, v in get_sorted(stats, time=t, type="residual_post_iteration", sortby="iter")] for t, _ in its]
    work = [int(v) for _, v in get_sorted(stats, type="work_CG", sortby="time")]
    return np.asarray(uend), [int(v) for _, v in its], res, work


def _worker(rank, world, port, name, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(HERE, "qmat_shim"))
    import pysdc_b200.mpi_facade

    sys.path.insert(0, pysdc_b200.mpi_facade.PATH)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mpi4py import MPI  # the transport shim

        spec = CASES[name]
        uend, niter, res, work = _run(spec, True, MPI.COMM_WORLD)
        out = dict(niter=niter, residuals=res, work_CG=work)
        if rank == 0:
            s_uend, s_niter, s_res, s_work = _run(spec, False, None)
            out.update(serial_niter=s_niter, serial_work_CG=s_work, serial_residuals=s_res,
                       serial_diff=float(np.max(np.abs(s_uend - uend)) / np.max(np.abs(s_uend))))
            np.save(os.path.join(out_dir, "serial_uend.npy"), s_uend)
        np.save(os.path.join(out_dir, f"uend_{rank}.npy"), uend)
        with open(os.path.join(out_dir, f"out_{rank}.json"), "w") as f:
            json.dump(out, f)
    finally:
        dist.destroy_process_group()


def main(which):
    for name in which:
        spec = CASES[name]
        world = spec["sweeper_params"]["num_nodes"]
        with tempfile.TemporaryDirectory() as tmp:
            mp.spawn(_worker, args=(world, free_port(), name, tmp), nprocs=world, join=True)
            outs = [json.load(open(os.path.join(tmp, f"out_{r}.json"))) for r in range(world)]
            uends = [np.load(os.path.join(tmp, f"uend_{r}.npy")) for r in range(world)]
        for r in range(1, world):  # every rank ends with the same uend, residual history and iteration counts
            assert np.array_equal(uends[r], uends[0]) and outs[r]["niter"] == outs[0]["niter"], (name, r)
            assert outs[r]["residuals"] == outs[0]["residuals"], (name, r)
        o = outs[0]
        assert o["niter"] == o["serial_niter"], (name, o["niter"], o["serial_niter"])
        assert o["serial_diff"] < 1e-10, (name, o["serial_diff"])
        width = max(len(r) for r in o["residuals"])
        res = np.full((len(o["residuals"]), width), np.nan)
        for i, r in enumerate(o["residuals"]):
            res[i, : len(r)] = r
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, spec=json.dumps(spec), niter=np.array(o["niter"]), residuals=res, uend=uends[0],
                            uend_maxabs=np.array(float(np.max(np.abs(uends[0])))),
                            work_CG_per_rank=np.array([x["work_CG"] for x in outs]),
                            serial_work_CG=np.array(o["serial_work_CG"]), serial_diff=np.array(o["serial_diff"]))
        print(f"wrote {path}: niter {o['niter']} (serial sweeper: {o['serial_niter']}, uend differs by {o['serial_diff']:.1e}) "
              f"CG per rank {[x['work_CG'] for x in outs]} serial {o['serial_work_CG']}")


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
