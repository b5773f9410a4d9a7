"""PFASST on GPUs against the reference's fixtures (run under torchrun, world size = num_procs of the fixture):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 \
        tests/mgpu/pfasst_check.py pfasst_heat2d_imex_63_p4 [nccl|gloo]

nccl: one GPU per rank, step-to-step hand-over by NCCL send/recv over NVLink.  gloo: all ranks share cuda:0 and device
fields are staged through the host (what the single-GPU test box can run)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    name = sys.argv[1]
    transport = sys.argv[2] if len(sys.argv) > 2 else "nccl"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if transport == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl")  # lazy init: point-to-point pairs get their own 2-rank communicators
    else:
        torch.cuda.set_device(0)
        dist.init_process_group("gloo")
    from conftest import load_golden
    from test_pfasst_gloo import _description
    from pysdc_b200.parallel import TorchComm
    from pysdc_b200.pfasst import controller_MPI
    from pysdc_b200.stats import get_sorted

    d, cp, t0, Tend, nprocs = _description(name)
    assert nprocs == world, f"fixture needs {nprocs} ranks"
    comm = TorchComm(device=torch.device("cuda", torch.cuda.current_device()) if transport == "nccl" else None)
    import time

    c = controller_MPI(cp, d, comm=comm)
    P = c.S.levels[0].prob
    u0 = P.u_exact(t0)
    uend, stats = c.run(u0=u0, t0=t0, Tend=Tend)  # also the warm-up of the timed run below
    torch.cuda.synchronize()
    dist.barrier()
    t_start = time.perf_counter()
    uend, stats = c.run(u0=u0, t0=t0, Tend=Tend)
    torch.cuda.synchronize()
    dist.barrier()
    wall = time.perf_counter() - t_start
    niter = [int(v) for _, v in get_sorted(stats, type="niter", sortby="time")]
    all_niter = [None] * world
    dist.all_gather_object(all_niter, niter)
    my_res = sorted((k.iter, float(v)) for k, v in stats.items() if k.type == "residual_post_iteration" and k.level == 0)
    all_res = [None] * world
    dist.all_gather_object(all_res, [v for _, v in my_res])
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        spec, g = load_golden(name)
        got = [v for r in all_niter for v in r]
        mine = uend.get()
        if "uend" in g:
            err = float(np.max(np.abs(mine - g["uend"])) / np.max(np.abs(g["uend"])))
        else:  # full-size fixture: subsample + max-norm of the reference's end value
            k = spec["subsample"]
            err = max(float(np.max(np.abs(mine[::k, ::k] - g["uend_sub"])) / float(g["uend_maxnorm"])),
                      abs(float(np.max(np.abs(mine))) - float(g["uend_maxnorm"])) / float(g["uend_maxnorm"]))
        ref = g["niter"].tolist()
        note = ""
        # Full-size run (BASELINE config 5): SDC stagnates at the accuracy the inner CG can attain, right at restol, and
        # the UNMODIFIED reference does not hold its own count on such slices: oracle/sensitivity.py re-ran it with
        # rounding-level perturbations of lintol / u0 and recorded where its count moves
        # (tests/golden/sensitivity_config5.json).  Only on those slices may a count differ, by one.
        import parity_cases as pc

        loose = pc.sensitive_steps(name, ref)
        ok_counts = len(got) == len(ref)
        for i, (a, b) in enumerate(zip(got, ref)):
            if a == b:
                continue
            ok_counts = ok_counts and i in loose and abs(a - b) == 1
            r_ref = float(g["residuals"][i][min(a, b) - 1]) if "residuals" in g else float("nan")
            note += (f" [slice {i}: {a} vs {b} iterations; the reference's residual at iteration {min(a, b)} is {r_ref:.3e}"
                     f" vs restol {d['level_params']['restol']:.0e}, ours {all_res[i][min(a, b) - 1]:.3e}; reference "
                     f"flips on this slice under rounding-level perturbations: {i in loose}]")
        ok = ok_counts and err < 1e-10
        nodes = c.S.levels[0].sweep.coll.num_nodes
        updates = P.dtype_u(P.init).size * nodes * sum(got)
        print(f"pfasst_check {name} world={world} {transport}: {'OK' if ok else 'FAILED'} niter={got} "
              f"(reference {g['niter'].tolist()}), uend rel. err {err:.2e}, wall {wall:.3f} s, "
              f"{updates / wall:.3e} fine-level DOF-node updates/s"
              + (f", reference CPU wall {float(g['wall_seconds']):.1f} s" if "wall_seconds" in g else "") + note, flush=True)
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
