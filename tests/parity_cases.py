"""Parity checks shared by the CPU suite (numpy test double for the kernel library: host logic only) and the GPU
suite (the real CUDA library through the C ABI).  Every check compares the package's classes with fixtures generated
by the unmodified reference (oracle/make_golden.py) and, where it needs fresh inputs, with the CPU oracle.

Tolerances (stated once, used everywhere):
  * eval_f, collocation integrals, end point: 1e-13 relative to the field's max-norm (pure stencil / axpy arithmetic,
    differences are summation order and FMA contraction only);
  * anything downstream of an iterative solve: 1e-10 relative (north_star: "end-of-step solution must agree to <= 1e-10
    relative error when both sides solve to the same lintol");
  * SDC iteration counts: identical.  CG / Newton work counters: identical up to +-1 per solve (the stopping test
    ||r|| < rtol*||b|| sits on a rounding-sensitive threshold), asserted as a relative band of 2 %.
"""
import numpy as np
import pytest

from conftest import load_golden

TOL_ARITH = 1e-13
TOL_SOLVE = 1e-10


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def stencil_tol(P, u_max, f_ref):
    """Absolute tolerance for f = A u: a few ulps of the largest term of the stencil sum (|a_diag| * max|u|), which
    for smooth fields is orders of magnitude larger than f itself (cancellation).  The one-sided closure rows of the
    higher-order Dirichlet stencils carry coefficients up to ~10x the centred diagonal."""
    if hasattr(P, "a_diag"):
        wide = 12.0 if getattr(P, "order", 2) != 2 else 1.0
        mag = wide * abs(P.a_diag)
    else:  # general operator tables (advection): the largest absolute row sum
        fd = P._fd
        rows = [np.abs(fd["coef"]).sum()] + ([] if fd["lo"] is None else [np.abs(fd["lo"]).sum(axis=1).max(),
                                                                         np.abs(fd["hi"]).sum(axis=1).max()])
        mag = P.ndim * max(rows)
    return 8 * np.finfo(float).eps * (mag * u_max + float(np.max(np.abs(f_ref))))


def classes():
    from pysdc_b200 import problems, sweepers

    return ({"heatNd_unforced": problems.heatNd_unforced, "heatNd_forced": problems.heatNd_forced,
             "advectionNd": problems.advectionNd,
             "allencahn_fullyimplicit": problems.allencahn_fullyimplicit,
             "allencahn_semiimplicit": problems.allencahn_semiimplicit,
             "allencahn_semiimplicit_v2": problems.allencahn_semiimplicit_v2,
             "allencahn_multiimplicit": problems.allencahn_multiimplicit,
             "allencahn_multiimplicit_v2": problems.allencahn_multiimplicit_v2},
            {"generic_implicit": sweepers.generic_implicit, "imex_1st_order": sweepers.imex_1st_order,
             "multi_implicit": sweepers.multi_implicit})


def tuplify(pp):
    pp = dict(pp)
    for k in ("nvars", "freq"):
        if isinstance(pp.get(k), list):
            pp[k] = tuple(pp[k])
    return pp


def make_description(spec):
    probs, sweeps = classes()
    return dict(problem_class=probs[spec["problem"]], problem_params=tuplify(spec["problem_params"]),
                sweeper_class=sweeps[spec["sweeper"]], sweeper_params=dict(spec["sweeper_params"]),
                level_params=dict(spec["level_params"]), step_params=dict(spec["step_params"]))


def to_mesh(P, arr, dtype=None):
    m = (dtype or P.dtype_u)(P.init)
    m[:] = arr
    return m


def close_counts(got, want, slack=0.02):
    got, want = np.atleast_1d(got), np.atleast_1d(want)
    return bool(np.all(np.abs(got - want) <= np.maximum(np.ceil(slack * want), 1)))


# ---------------------------------------------------------------------------------------------------------------------
def check_operator(name):
    spec, g = load_golden(name)
    probs, _ = classes()
    P = probs[spec["problem"]](**tuplify(spec["problem_params"]))
    u = to_mesh(P, g["u"])
    rhs = to_mesh(P, g["rhs"])
    u_before, rhs_before = u.get().copy(), rhs.get().copy()
    f = P.eval_f(u, spec["t"])
    assert type(f) is P.dtype_f
    assert f.shape == g["f"].shape
    assert np.max(np.abs(f.get() - g["f"])) <= stencil_tol(P, np.max(np.abs(g["u"])), g["f"])
    if "sol1" in g:  # the two solves of the multi-implicit splitting (AllenCahn_2D_FD.py:534-651,699-776)
        sol1 = P.solve_system_1(rhs, spec["factor"], u, spec["t"])
        assert type(sol1) is P.dtype_u and relerr(sol1.get(), g["sol1"]) < TOL_SOLVE
        assert P.newton_itercount == int(g["newton_after_1"]) and close_counts(P.lin_itercount, int(g["linear_after_1"]))
        sol2 = P.solve_system_2(rhs, spec["factor"], u, spec["t"])
        assert type(sol2) is P.dtype_u and relerr(sol2.get(), g["sol2"]) < TOL_SOLVE
        assert P.newton_itercount == int(g["newton_itercount"]) and close_counts(P.lin_itercount, int(g["lin_itercount"]))
        assert np.array_equal(u.get(), u_before) and np.array_equal(rhs.get(), rhs_before)
        assert all(c.niter == 0 for c in P.work_counters.values())  # these classes count in the plain ints only
        with pytest.raises(Exception):
            P.solve_system(rhs, spec["factor"], u, spec["t"])
        return
    sol = P.solve_system(rhs, spec["factor"], u, spec["t"])
    assert type(sol) is P.dtype_u
    assert relerr(sol.get(), g["sol"]) < TOL_SOLVE
    # inputs are not mutated (generic_ND_FD.py:208-264 returns a fresh field)
    assert np.array_equal(u.get(), u_before) and np.array_equal(rhs.get(), rhs_before)
    if "cg_iters" in g:
        assert close_counts(P.work_counters["CG"].niter, int(g["cg_iters"]))
    elif "gmres_iters" in g:  # inner iterations = calls of the reference's work counter (callback_type='legacy')
        assert close_counts(P.work_counters["GMRES"].niter, int(g["gmres_iters"]))
    else:
        assert P.work_counters["newton"].niter == int(g.get("newton", 0))
        assert close_counts(P.work_counters["linear"].niter, int(g["linear"]))
        v2 = spec["problem"].endswith("_v2")  # counts Newton steps in newton_itercount only, eval_f not at all
        assert P.work_counters["rhs"].niter == (0 if v2 else 1)
        if "newton_itercount" in g:
            assert P.newton_itercount == int(g["newton_itercount"]) and P.newton_ncalls == 1
    t_ex = 0.0 if spec["problem"].startswith("allencahn") else 0.1
    assert relerr(P.u_exact(t_ex).get(), g["u_exact"]) == 0.0  # host numpy expression, then upload


def check_sweep_dump(name):
    from pysdc_b200.core import Step

    spec, g = load_golden(name)
    S = Step(make_description(spec))
    L = S.levels[0]
    P = L.prob
    L.status.time = spec["t0"]
    np.testing.assert_allclose(L.sweep.coll.Qmat, g["Qmat"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(L.sweep.coll.nodes, g["nodes"], rtol=0, atol=1e-15)
    S.init_step(to_mesh(P, g["u0"]))
    L.sweep.predict()
    assert L.status.unlocked and L.status.updated
    if "tau" in g:
        L.tau = [to_mesh(P, t) for t in g["tau"]]
    f_pred = np.stack([f.get() for f in L.f])
    assert np.max(np.abs(f_pred - g["f_pred"])) <= stencil_tol(P, np.max(np.abs(g["u0"])), g["f_pred"])
    L.sweep.compute_residual()
    assert isinstance(L.status.residual, float) and not L.status.updated
    assert abs(L.status.residual - float(g["res_pred"])) <= 1e-12 * max(1.0, abs(float(g["res_pred"])))
    k = 1
    while f"u_sweep{k}" in g:
        L.status.sweep = k
        L.sweep.updateVariableCoeffs(k)
        u_ids = [id(u) for u in L.u]
        L.sweep.update_nodes()
        assert L.status.updated
        assert relerr(np.stack([u.get() for u in L.u]), g[f"u_sweep{k}"]) < TOL_SOLVE, k
        assert relerr(np.stack([f.get() for f in L.f]), g[f"f_sweep{k}"]) < 10 * TOL_SOLVE, k
        integ = L.sweep.integrate()
        assert len(integ) == L.sweep.coll.num_nodes and all(type(i) is P.dtype_u for i in integ)
        assert relerr(np.stack([i.get() for i in integ]), g[f"integrate_sweep{k}"]) < 10 * TOL_SOLVE
        for rt in ("full_abs", "last_abs", "full_rel", "last_rel"):
            L.params.residual_type = rt
            L.sweep.compute_residual()
            want = float(g[f"res_{rt}_sweep{k}"])
            # the residual is a difference of O(1) quantities each carrying the 1e-10 solve tolerance
            assert abs(L.status.residual - want) <= 1e-9 * max(abs(float(g["res_pred"])), 1.0), (rt, k)
        L.params.residual_type = "full_abs"
        L.sweep.compute_residual()
        # L.residual as the reference leaves it (core/sweeper.py:188-193): res[m] = integrate()[m] + u0 - u[m+1] (+ tau)
        assert len(L.residual) == L.sweep.coll.num_nodes
        for m, (res, q) in enumerate(zip(L.residual, integ)):
            want = q + L.u[0] - L.u[m + 1]
            if L.tau[m] is not None:
                want += L.tau[m]
            assert type(res) is P.dtype_u and abs(res - want) <= 1e-12 * max(1.0, abs(L.u[0]))
        assert abs(max(abs(r) for r in L.residual) - L.status.residual) <= 1e-14 * max(1.0, L.status.residual)
        assert len(u_ids) == len(L.u)
        k += 1
    np.testing.assert_allclose(L.sweep.QI, g["QI"], rtol=0, atol=1e-14)
    L.sweep.compute_end_point()
    assert type(L.uend) is P.dtype_u
    assert relerr(L.uend.get(), g["uend"]) < TOL_SOLVE
    for key in P.work_counters:
        want = int(g["work_" + key])
        if key in ("newton", "rhs"):
            assert P.work_counters[key].niter == want, key
        else:
            assert close_counts(P.work_counters[key].niter, want), (key, P.work_counters[key].niter, want)


def check_spatial_accuracy(pmax):
    """The reference's own pin for the higher-order stencils (pySDC/tests/test_2d_fd_accuracy.py:10-34): the error of
    eval_f on a sine wave against the analytic Laplacian shrinks with the stencil's order (2, 4, 8) on periodic 2-D
    grids of 2^4 .. 2^pmax points per dimension (errors below 1e-8 are left out, as in the reference's test)."""
    probs, _ = classes()
    for order_stencil in (2, 4, 8):
        errs, sizes = [], []
        for p in range(4, pmax + 1):
            n = 2**p
            P = probs["heatNd_unforced"](nvars=(n, n), freq=(2, 2), nu=1.0, bc="periodic", order=order_stencil,
                                          solver_type="CG")
            x = np.array([i * P.dx for i in range(n)])
            u_lap = to_mesh(P, -2 * (np.pi**2 * P.freq[0] * P.freq[1]) * P.nu
                            * np.kron(np.sin(np.pi * P.freq[0] * x), np.sin(np.pi * P.freq[1] * x)).reshape(n, n))
            errs.append(abs(P.eval_f(P.u_exact(0.0), 0.0) - u_lap))
            sizes.append(n)
        order = [np.log(errs[i - 1] / errs[i]) / np.log(sizes[i] / sizes[i - 1]) for i in range(1, len(errs))
                 if errs[i] > 1e-8 and errs[i - 1] > 1e-8]
        assert len(order) >= 1 and np.allclose(order, order_stencil, atol=5e-2 if pmax >= 10 else 0.35), (order_stencil, order)


def sensitive_steps(name, ref_niter):
    """Steps of a BASELINE-size fixture on which the UNMODIFIED reference does not hold its own SDC iteration count: runs
    of the reference with rounding-level perturbations of its inputs, or with another BLAS thread count, recorded in
    tests/golden/sensitivity_config*.json by oracle/sensitivity.py.  Only there may a count differ (by one)."""
    import json
    import os

    from conftest import GOLDEN

    cfg = {"run_config2_heat2d_imex_lu_2047": "config2", "pfasst_config5_1023_p8": "config5"}.get(name)
    path = os.path.join(GOLDEN, f"sensitivity_{cfg}.json")
    if cfg is None or not os.path.exists(path):
        return set()
    with open(path) as f:
        runs = json.load(f)["runs"]
    return {i for r in runs for i, (a, b) in enumerate(zip(r["niter"], ref_niter)) if a != b}


def check_run(name, uend_tol=TOL_SOLVE, count_slack=0.02):
    from pysdc_b200.controller import LogWork, controller_nonMPI
    from pysdc_b200.stats import get_sorted

    spec, g = load_golden(name)
    d = make_description(spec)
    c = controller_nonMPI(num_procs=1, controller_params={"logger_level": 40, "hook_class": [LogWork]}, description=d)
    P = c.MS[0].levels[0].prob
    if spec["u0"] == "exact":
        u0 = P.u_exact(spec["t0"])
    else:
        u0 = to_mesh(P, np.random.default_rng(spec["seed"]).standard_normal(P.nvars))
    uend, stats = c.run(u0=u0, t0=spec["t0"], Tend=spec["Tend"])
    niter = [int(v) for _, v in get_sorted(stats, type="niter", sortby="time")]
    want = g["niter"].tolist()
    loose = sensitive_steps(name, want)  # empty for every fixture but the rounding-decided BASELINE-size ones
    assert len(niter) == len(want) and all(a == b or (i in loose and abs(a - b) == 1)
                                           for i, (a, b) in enumerate(zip(niter, want))), (niter, want, loose)
    times = [t for t, _ in get_sorted(stats, type="niter", sortby="time")]
    for i, (t, ref) in enumerate(zip(times, g["residuals"])):
        hist = [v for _, v in get_sorted(stats, time=t, type="residual_post_iteration", sortby="iter")]
        ref = ref[~np.isnan(ref)]
        assert len(hist) == len(ref) or i in loose
        k = min(len(hist), len(ref))
        # per-sweep residual histories agree to 1e-8 relative above the solver noise floor (the
        # BASELINE-size runs stagnate at ~1e-10, where the inner CG's rounding shows)
        floor = (6e-11 if name.startswith("run_config") else 2e-11) * max(1.0, float(g["uend_maxabs"]))
        np.testing.assert_allclose(hist[:k], ref[:k], rtol=1e-6, atol=floor)
    # error relative to the solution scale of the run (the initial value for strongly decaying solutions: the SDC
    # residual tolerance that terminates every step is absolute)
    scale = max(abs(u0), float(g["uend_maxabs"]))
    if "uend" in g:
        assert np.max(np.abs(uend.get() - g["uend"])) <= uend_tol * scale
    if "uend_sub" in g:  # BASELINE-size fixtures keep a subsample of the end value
        k = spec["subsample"]
        sub = uend.get()[(slice(None, None, k),) * uend.get().ndim]
        assert np.max(np.abs(sub - g["uend_sub"])) <= uend_tol * scale
    assert abs(abs(uend) - float(g["uend_maxabs"])) <= uend_tol * scale
    for key in P.work_counters:
        got = [int(v) for _, v in get_sorted(stats, type="work_" + key, sortby="time")]
        want = g["work_" + key].tolist()
        if key in ("newton", "rhs"):
            assert got == want, (key, got, want)
        elif count_slack is not None:
            # BASELINE-size runs: thousands of CG iterations per step, each solve's count within a few of the reference's;
            # a step that takes one sweep more or less (see sensitive_steps) is not comparable
            keep = [i for i in range(len(want)) if i not in loose or niter[i] == g["niter"][i]]
            slack = 0.06 if name.startswith("run_config") else count_slack
            assert close_counts([got[i] for i in keep], [want[i] for i in keep], slack), (key, got, want)
    if "newton_itercount" in g:  # the reference's plain-int totals (AllenCahn_2D_FD.py:202-203,481-482)
        assert P.newton_itercount == int(g["newton_itercount"]) and P.newton_ncalls == int(g["newton_ncalls"])
    if "lin_itercount" in g:
        assert close_counts(P.lin_itercount, int(g["lin_itercount"]), count_slack or 0.02) and P.lin_ncalls == int(g["lin_ncalls"])
    return dict(niter=niter, uend=uend, stats=stats)


# ---------------------------------------------------------------------------------------------------------------------
def check_transfer(name):
    """mesh_to_mesh restriction / prolongation vs the reference's sparse Kronecker-product operators
    (transfer_classes/TransferMesh.py:149-218) on the fixture's seeded fields."""
    from pysdc_b200.transfer import mesh_to_mesh

    spec, g = load_golden(name)
    probs, _ = classes()
    pp = tuplify(spec["problem_params"])
    as_nvars = lambda v: tuple(v) if isinstance(v, list) else v  # noqa: E731
    Pf = probs[spec["problem"]](nvars=as_nvars(spec["nvars_fine"]), solver_type="CG", **pp)
    Pc = probs[spec["problem"]](nvars=as_nvars(spec["nvars_coarse"]), solver_type="CG", **pp)
    T = mesh_to_mesh(Pf, Pc, dict(spec["transfer_params"]))
    F, G = to_mesh(Pf, g["F"]), to_mesh(Pc, g["G"])
    RF, PG = T.restrict(F), T.prolong(G)
    assert type(RF) is Pc.dtype_u and type(PG) is Pf.dtype_u
    assert RF.shape == g["RF"].shape and PG.shape == g["PG"].shape
    assert relerr(RF.get(), g["RF"]) < TOL_ARITH and relerr(PG.get(), g["PG"]) < TOL_ARITH
    assert np.array_equal(F.get(), g["F"]) and np.array_equal(G.get(), g["G"])  # inputs untouched
    if "Ff" in g:
        Ff, Gf = to_mesh(Pf, g["Ff"], Pf.dtype_f), to_mesh(Pc, g["Gf"], Pc.dtype_f)
        RFf, PGf = T.restrict(Ff), T.prolong(Gf)
        assert type(RFf) is Pc.dtype_f and type(PGf) is Pf.dtype_f
        assert relerr(RFf.get(), g["RFf"]) < TOL_ARITH and relerr(PGf.get(), g["PGf"]) < TOL_ARITH
