#!/usr/bin/env python
"""Headline benchmark: SDC sweep DOF-node updates/s on the configuration BASELINE.json quotes
(3-D heat 511^3 fp64, generic_implicit, M=4 RADAU-RIGHT nodes, QI='MIN-SR-NS' -> node-batched solves).

    python bench.py --gpus 1 --steps 3 --warmup 3                 # this repo's CUDA path
    python bench.py --impl reference --steps 1 --warmup 0         # the reference algorithm on the host cores

One "step" = one SDC time step on a fresh seeded random field: predict + K=4 x (update_nodes + compute_residual) +
compute_end_point, i.e. N*M*4 DOF-node updates.  Prints ONE JSON line (contract in the task description; field meanings
in DESIGN.md section "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_NODES, K_SWEEPS = 4, 4
METRIC = "SDC sweep DOF-node updates/s (3D heat 511^3, M=4 MIN-SR-NS, fp64)"
UNIT = "DOF-node updates/s"


def spec_for(n, K=K_SWEEPS):
    """Description of the headline workload for an n^3 grid, as a fixture-style spec."""
    return dict(problem="heatNd_unforced", sweeper="generic_implicit",
                problem_params=dict(nvars=[n, n, n], nu=0.1, freq=[1, 1, 1], bc="dirichlet-zero", solver_type="CG",
                                    lintol=1e-12, liniter=10000),
                sweeper_params=dict(num_nodes=M_NODES, quad_type="RADAU-RIGHT", QI="MIN-SR-NS", initial_guess="spread"),
                level_params=dict(dt=1e-3, restol=-1.0), step_params=dict(maxiter=K),
                t0=0.0, Tend=1e-3, u0="random", seed=1234)


def measured_traffic(n, world):
    """Mean DRAM bytes per CG launch from the committed ncu capture of this very command (profiles/): only meaningful
    for the configuration it was taken on (511^3, one GPU)."""
    try:
        if n != 511 or world != 1:
            return None, None
        with open(os.path.join(ROOT, "profiles", "cg_traffic_511.json")) as f:
            t = json.load(f)
        per = [l["traffic"] for l in t["launches"]]
        return float(np.mean(per)), "profiles/cg_traffic_511.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean of 8 launches)"
    except Exception:
        return None, None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle restatement of the reference algorithm (scipy sparse + scipy cg)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(n, K):
    """One SDC step of the headline description at n^3 with the oracle port; the sparse operator is assembled outside
    the timed region (the reference does that in the problem constructor)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sdc_oracle

    spec = spec_for(n, K)
    t0 = time.perf_counter()
    L = sdc_oracle.make_level(spec)
    t_setup = time.perf_counter() - t0
    u0 = sdc_oracle.initial_value(L.prob, spec)
    t0 = time.perf_counter()
    out = sdc_oracle.run_sdc(spec, u0=u0, level=L)
    secs = time.perf_counter() - t0
    return dict(rate=n**3 * M_NODES * K / secs, seconds=secs, setup_seconds=t_setup, n=n, K=K,
                cg_per_solve=out["work"]["CG"][0] / (M_NODES * K))


def run_reference(args):
    n = args.ref_n
    t_all = time.perf_counter()
    rates = []
    # bounded: every step is ~10-25 s of single-core work; at most one untimed warm-up step and ~2.5 minutes in total
    for _ in range(min(args.warmup, 1)):
        cpu_reference_rate(n, K_SWEEPS)
    for _ in range(max(args.steps, 1)):
        rates.append(cpu_reference_rate(n, K_SWEEPS))
        if time.perf_counter() - t_all > 150.0:
            break
    secs = sum(r["seconds"] for r in rates)
    value = n**3 * M_NODES * K_SWEEPS * len(rates) / secs
    sample = (f"oracle port of the reference path (scipy sparse matvec + scipy cg), same description at {n}^3 "
              f"(bounded sample of the 511^3 workload), {K_SWEEPS} sweeps/step, {rates[0]['cg_per_solve']:.1f} CG it/solve")
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=len(rates),
                warmup=args.warmup, ms_per_step=1e3 * secs / len(rates), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=f"heat3d_{n}cubed_M4_MINSRNS_K{K_SWEEPS} (CPU sample of heat3d_511cubed)",
                            inputs="seeded N(0,1) field, default_rng(1234)"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=1, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                wall_s=time.perf_counter() - t_all)
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from pysdc_b200 import backend as bk
    from pysdc_b200.controller import controller_nonMPI
    from pysdc_b200.problems import heatNd_unforced
    from pysdc_b200.sweepers import generic_implicit

    be = bk.get_backend()
    n = args.n
    spec = spec_for(n)
    pp = dict(spec["problem_params"])
    pp["nvars"], pp["freq"] = tuple(pp["nvars"]), tuple(pp["freq"])
    if args.precond:
        pp["preconditioner"] = "chebyshev"
    comm = None
    if world > 1:
        # the SAME n^3 problem, slab-decomposed along axis 0 over the GPUs of the node ("strong" scaling)
        from pysdc_b200.parallel import SlabComm

        comm = SlabComm()
        pp["comm"] = comm
    description = dict(problem_class=heatNd_unforced, problem_params=pp, sweeper_class=generic_implicit,
                       sweeper_params=dict(spec["sweeper_params"]), level_params=dict(spec["level_params"]),
                       step_params=dict(spec["step_params"]))
    ctrl = controller_nonMPI(num_procs=1, controller_params={"logger_level": 40}, description=description)
    P = ctrl.MS[0].levels[0].prob

    # seeded N(0,1) field of the global grid, generated on the host plane by plane; a rank keeps the planes it owns
    lay = P._lay
    nz, z0 = (lay.nz, lay.z0) if lay.is_slab else (n, 0)
    rng = np.random.default_rng(1234)
    host_u0 = torch.empty((nz, n, n), dtype=torch.float64).pin_memory()
    for z in range(z0 + nz):
        plane = rng.standard_normal((n, n))
        if z >= z0:
            host_u0.numpy()[z - z0] = plane
    host_uend = torch.empty((nz, n, n), dtype=torch.float64).pin_memory()
    u0 = P.dtype_u(P.init)
    u0.data.copy_(host_u0, non_blocking=True)
    dof_updates_per_step = n**3 * M_NODES * K_SWEEPS  # of the whole job, whatever the number of GPUs

    def step():
        return ctrl.run(u0=u0, t0=0.0, Tend=1e-3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    # ---- device-resident timing -------------------------------------------------------------------------------------
    P.solve_log = []
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = be.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        uend, stats = step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = be.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    solve_log, P.solve_log = P.solve_log, None

    # ---- end to end through the public API with host buffers --------------------------------------------------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    copy_stream = torch.cuda.Stream()
    for _ in range(args.steps):
        u0.data.copy_(host_u0, non_blocking=True)           # H2D of the step's input from pinned memory
        uend, stats = step()
        # D2H of the step's result on a side stream: it overlaps the next step's compute (uend is a fresh buffer per
        # step); the timed region ends only after the last copy has landed
        copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):
            host_uend.copy_(uend.data, non_blocking=True)
            uend.data.record_stream(copy_stream)
    torch.cuda.current_stream().wait_stream(copy_stream)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    # ---- the streaming kernels of the sweep on their own (north_star: "collocation integration ... one coalesced
    # kernel fused with the rhs assembly and the residual norm"): algorithmic bytes / CUDA-event time ----------------
    other = {}
    if rank == 0 and world == 1:
        L = ctrl.MS[0].levels[0]
        sweep = L.sweep
        nloc = nz * n * n
        reps = 10

        def timed(fn):
            fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        fins = sweep._f_inputs(L)
        rhs = sweep._scratch(L, M_NODES)
        W = np.random.default_rng(0).standard_normal((M_NODES, len(fins)))
        norms = be.zeros(M_NODES)
        kernels = {
            "colloc_sweep_kernel<4,1> (rhs assembly: u0 + dt(Q-QD)F)":
                (lambda: be.colloc_sweep(fins, 1, [r.flat for r in rhs], Wq=W, Wi=-W, base=L.u[0].flat), 8 * (2 * M_NODES + 1)),
            "colloc_residual_kernel<4,1> (residual + max-norms)":
                (lambda: be.colloc_residual(W, fins, 1, L.u[0].flat, [u.flat for u in L.u[1:]], None, None, norms),
                 8 * (2 * M_NODES + 1)),
            "eval_f_kernel<3> (4 fields)":
                (lambda: P.eval_f_batch(L.u[1:], [0.0] * M_NODES, L.f[1:]), 16 * M_NODES),
        }
        for name, (fn, bytes_per_dof) in kernels.items():
            t_ms = timed(fn)
            gbs = bytes_per_dof * nloc / (t_ms * 1e-3) / 1e9
            other[name] = dict(ms=t_ms, algorithmic_bytes=bytes_per_dof * nloc, achieved_gbs=gbs)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = (float(v) for v in t.tolist())

    if rank == 0:
        peak, peak_src = peaks()
        # dominant kernel: the persistent batched CG.  Algorithmic bytes per launch:
        # sum over the B systems of 8 B * N * (4 [set-up: read b, x0; write r, p] + 9 * iterations)
        cg_ms = sum(a.elapsed_time(b) for a, b, _ in solve_log)
        cg_iters = torch.stack([c for _, _, c in solve_log]).cpu().numpy().astype(np.int64)
        # plain CG: 4 streams of set-up + 9 per iteration; the polynomial preconditioner adds one pass (read r, write z)
        per_it, setup = (11, 6) if args.precond else (9, 4)
        alg_bytes = 8.0 * nz * n * n * float(np.sum(setup + per_it * cg_iters))  # this rank's share
        n_launch = len(solve_log)
        achieved = alg_bytes / (cg_ms * 1e-3) / 1e9
        n_cg = float(cg_iters.mean())
        b_alg = 84 + (16 + 88 * n_cg if args.precond else 72 * n_cg)
        traffic, traffic_src = (None, None) if args.precond else measured_traffic(n, world)
        value = dof_updates_per_step * args.steps / (ms * 1e-3)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="strong" if world > 1 else "weak",
                    vs_baseline=None,
                    dtype="f64", data="synthetic",
                    config=dict(workload=f"heat3d_{n}cubed_M4_MINSRNS_K{K_SWEEPS}", parallelism=f"{world} slabs along axis 0 (peer-memory CG)" if world > 1 else "single GPU",
                                cache="working set per step ~28 GB >> 126 MB L2: no flush needed",
                                inputs="seeded N(0,1) field, default_rng(1234)", cg_it_per_solve=n_cg,
                                solver=("CG preconditioned with a degree-1 Chebyshev polynomial of the operator "
                                        "(same lintol and stopping test as the reference's plain CG)") if args.precond
                                else "plain CG (the reference's algorithm)",
                                b_alg_bytes_per_update=b_alg),
                    e2e=dict(value=dof_updates_per_step * args.steps / (ms_e2e * 1e-3), unit=UNIT,
                             h2d_bytes_per_step=8 * n**3, d2h_bytes_per_step=8 * n**3),
                    gpu_launches=launches,
                    roofline=dict(bound="hbm", kernel="cg_pipe_kernel<3> (persistent node-batched CG, TMA-pipelined passes)"
                                  + (", rank 0's slab" if world > 1 else ""), achieved=achieved,
                                  peak=peak, unit="GB/s", frac=achieved / peak, peak_source=peak_src, traffic=traffic,
                                  traffic_unit="bytes per launch", traffic_source=traffic_src,
                                  algorithmic_bytes_per_launch=alg_bytes / max(n_launch, 1),
                                  launches=n_launch, ms_per_launch=cg_ms / max(n_launch, 1),
                                  share_of_step=cg_ms / ms,
                                  whole_step_achieved=b_alg * dof_updates_per_step * args.steps / (ms * 1e-3) / 1e9),
                    other_kernels={k: dict(v, frac_of_peak=v["achieved_gbs"] / peak) for k, v in other.items()},
                    clocks=clocks)
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_rate(args.ref_n, K_SWEEPS)
            line["cpu_baseline"] = dict(
                value=r["rate"], unit=UNIT, cores=1, kind="port",
                sample=(f"oracle port of the reference path (scipy sparse matvec + scipy cg, single-threaded) on this "
                        f"host, same description at {r['n']}^3, 1 step of {K_SWEEPS} sweeps in {r['seconds']:.1f} s, "
                        f"{r['cg_per_solve']:.1f} CG it/solve; host has {os.cpu_count()} cores"))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=511, help="grid points per dimension (headline: 511)")
    ap.add_argument("--ref-n", type=int, default=127, help="grid size of the bounded CPU sample (127^3: ~15 s per step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precond", action="store_true", help="node solves with the polynomial preconditioner")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", 0)) == 0:
            run_reference(args)
        return
    run_b200(args)


if __name__ == "__main__":
    main()
