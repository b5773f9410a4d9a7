#!/usr/bin/env python
"""Device timings (CUDA events, after warm-up) of the kernels outside the headline step at sizes that leave the L2:
restarted GMRES (advection 3-D), order-4 CG (heat 3-D), the point-wise reaction Newton (Allen-Cahn 4096^2), with their
algorithmic bytes against the measured HBM peak.  One JSON line per kernel on stdout.

    python scripts/profile_extra.py [gmres|ho_cg|reaction ...]      (under ncu: one kernel each, see scripts/gpu_r2v.sh)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pysdc_b200 import backend  # noqa: E402
from pysdc_b200 import problems  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]


def timed(fn, reps=3):
    fn()  # warm-up
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def mesh_of(P, arr):
    m = P.dtype_u(P.init)
    m[:] = arr
    return m


def gmres():
    n = 256
    P = problems.advectionNd(nvars=(n, n, n), c=1.0, freq=(2, 2, 2), stencil_type="upwind", order=3, bc="periodic",
                             solver_type="GMRES", lintol=1e-11, liniter=60)
    rng = np.random.default_rng(1)
    rhs, x0 = mesh_of(P, rng.standard_normal(P.nvars)), mesh_of(P, rng.standard_normal(P.nvars))
    x = P.dtype_u(x0)
    before = [0]

    def run():
        x[:] = x0
        before[0] = P.work_counters["GMRES"].niter
        P.solve_system_batch([rhs], [2e-3], [x], [0.0])

    ms = timed(run)
    its = P.work_counters["GMRES"].niter - before[0]
    # inner iteration k of a restart cycle (k = 0..19): operator 16 B + Gram-Schmidt 24 B * (k + 1) per DOF; + the cycle's
    # residual / update passes (~ 2 operator passes + 20 axpys)
    N = n**3
    per_cycle = sum(16 + 24 * (k + 1) for k in range(20)) + 2 * 16 + 20 * 16
    alg = N * per_cycle * (its / 20.0)
    return dict(kernel="gmres_kernel", workload=f"advectionNd {n}^3 upwind order 3 periodic, {its} inner iterations",
                ms=ms, algorithmic_bytes=alg, achieved_gbs=alg / ms / 1e6, frac_of_measured_peak=alg / ms / 1e6 / PEAK)


def ho_cg():
    n = 255
    P = problems.heatNd_unforced(nvars=(n, n, n), nu=0.1, freq=(1, 1, 1), bc="dirichlet-zero", order=4, solver_type="CG",
                                 lintol=1e-12, liniter=10000)
    rng = np.random.default_rng(2)
    rhs, x0 = mesh_of(P, rng.standard_normal(P.nvars)), mesh_of(P, rng.standard_normal(P.nvars))
    x = P.dtype_u(x0)
    before = [0]

    def run():
        x[:] = x0
        before[0] = P.work_counters["CG"].niter
        P.solve_system_batch([rhs], [2e-4], [x], [0.0])

    ms = timed(run)
    its = P.work_counters["CG"].niter - before[0]
    alg = n**3 * (32 + 88.0 * its)
    return dict(kernel="ho_cg_kernel", workload=f"heatNd_unforced {n}^3 order 4 dirichlet-zero, {its} CG iterations", ms=ms,
                algorithmic_bytes=alg, achieved_gbs=alg / ms / 1e6, frac_of_measured_peak=alg / ms / 1e6 / PEAK)


def reaction():
    n = 4096
    P = problems.allencahn_multiimplicit(nvars=(n, n), nu=2, eps=0.04, newton_maxiter=100, newton_tol=1e-9, lin_tol=1e-10,
                                         lin_maxiter=100, radius=0.25)
    u0 = P.u_exact(0.0)
    rng = np.random.default_rng(3)
    rhs = mesh_of(P, u0.get() + 0.05 * rng.standard_normal(P.nvars))
    x = P.dtype_u(u0)
    before = [0]

    def run():
        x[:] = u0
        before[0] = P.newton_itercount
        P.solve_system_2_batch([rhs], [1e-3], [x], [0.0])

    ms = timed(run)
    its = P.newton_itercount - before[0]
    alg = n * n * (16 + 24.0 * its)  # first pass reads u, rhs; every update reads u, rhs and writes u
    return dict(kernel="reaction_newton_kernel", workload=f"allencahn_multiimplicit.solve_system_2 {n}^2, {its} Newton updates",
                ms=ms, algorithmic_bytes=alg, achieved_gbs=alg / ms / 1e6, frac_of_measured_peak=alg / ms / 1e6 / PEAK)


if __name__ == "__main__":
    backend.set_backend(backend.CudaBackend())
    for name in sys.argv[1:] or ["gmres", "ho_cg", "reaction"]:
        print(json.dumps(dict(globals()[name](), peak_gbs=PEAK)), flush=True)
