"""Pin the CPU oracle (oracle/sdc_oracle.py) against fixtures produced by the unmodified reference
(oracle/make_golden.py).  CPU only; this is what lets the GPU parity tests trust the oracle."""
import numpy as np
import pytest

from conftest import golden_names, load_golden


def _relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("name", golden_names("op_"))
def test_operator_vectors(oracle, name):
    spec, g = load_golden(name)
    P = oracle.make_problem(spec["problem"], spec["problem_params"])
    f = P.eval_f(g["u"], spec["t"])
    assert _relerr(f, g["f"]) == 0.0  # same scipy matvec, same expression order
    if "sol1" in g:  # the two solves of the multi-implicit splitting (AllenCahn_2D_FD.py:534-651,699-776)
        assert _relerr(P.solve_system_1(g["rhs"], spec["factor"], g["u"], spec["t"]), g["sol1"]) == 0.0
        assert (P.newton_itercount, P.lin_itercount) == (int(g["newton_after_1"]), int(g["linear_after_1"]))
        assert _relerr(P.solve_system_2(g["rhs"], spec["factor"], g["u"], spec["t"]), g["sol2"]) == 0.0
        assert (P.newton_itercount, P.lin_itercount) == (int(g["newton_itercount"]), int(g["lin_itercount"]))
        return
    sol = P.solve_system(g["rhs"], spec["factor"], g["u"], spec["t"])
    assert _relerr(sol, g["sol"]) == 0.0
    if "cg_iters" in g:
        assert P.counters["CG"].niter == int(g["cg_iters"])
    elif "gmres_iters" in g:
        assert P.counters["GMRES"].niter == int(g["gmres_iters"])
    else:
        assert P.counters["newton"].niter == int(g["newton"]) if "newton" in g else P.counters["newton"].niter == 0
        assert P.counters["linear"].niter == int(g["linear"])
        if "newton_itercount" in g:
            assert P.newton_itercount == int(g["newton_itercount"])
    t_ex = 0.0 if spec["problem"].startswith("allencahn") else 0.1
    assert _relerr(P.u_exact(t_ex), g["u_exact"]) == 0.0


@pytest.mark.parametrize("name", golden_names("sweep_"))
def test_sweep_dumps(oracle, name):
    spec, g = load_golden(name)
    L = oracle.make_level(spec)
    L.time = spec["t0"]
    if not L.genQI.isKDependent():
        np.testing.assert_array_equal(L.QI, g["QI"])
    np.testing.assert_array_equal(L.coll.Qmat, g["Qmat"])
    L.u[0] = g["u0"].copy()
    oracle.predict(L)
    if "tau" in g:
        L.tau = [t.copy() for t in g["tau"]]
    assert _relerr(np.stack(L.f), g["f_pred"]) == 0.0
    assert oracle.compute_residual(L) == float(g["res_pred"])
    k = 1
    while f"u_sweep{k}" in g:
        L.update_variable_coeffs(k)
        oracle.update_nodes(L)
        assert _relerr(np.stack(L.u), g[f"u_sweep{k}"]) < 1e-15
        assert _relerr(np.stack(L.f), g[f"f_sweep{k}"]) < 1e-13
        assert _relerr(np.stack(oracle.integrate(L)), g[f"integrate_sweep{k}"]) < 1e-13
        for rt in ("full_abs", "last_abs", "full_rel", "last_rel"):
            L.residual_type = rt
            assert oracle.compute_residual(L) == pytest.approx(float(g[f"res_{rt}_sweep{k}"]), rel=1e-12)
        L.residual_type = "full_abs"
        oracle.compute_residual(L)
        k += 1
    assert _relerr(oracle.compute_end_point(L), g["uend"]) < 1e-15
    np.testing.assert_array_equal(L.QI, g["QI"])  # k-dependent generators: coefficients of the last sweep
    for key, c in L.prob.counters.items():
        assert c.niter == int(g["work_" + key]), key


# of the BASELINE-size fixtures (run_config*) the oracle replays the two it finishes in seconds
@pytest.mark.parametrize("name", [n for n in golden_names("run_") if "63_K4" not in n and "255" not in n
                                  and (not n.startswith("run_config") or n.endswith(("_511", "_256", "_127_K4")))])
def test_full_runs(oracle, name):
    spec, g = load_golden(name)
    out = oracle.run_sdc(spec)
    assert out["niter"] == g["niter"].tolist()
    for key in out["work"]:
        assert out["work"][key] == g["work_" + key].tolist(), key
    for hist, ref in zip(out["residuals"], g["residuals"]):
        ref = ref[~np.isnan(ref)]
        # (atol: numpy's dot may split long vectors over BLAS threads, which moves the CG iterates by an ulp or two)
        np.testing.assert_allclose(hist, ref, rtol=1e-9, atol=1e-14)
    if "uend" in g:
        assert _relerr(out["uend"], g["uend"]) < 1e-14
    else:
        sub = spec["subsample"]
        assert _relerr(out["uend"][(slice(None, None, sub),) * out["uend"].ndim], g["uend_sub"]) < 1e-13


def test_reference_known_answers():
    """The reference's own pins, carried by the fixtures (asserted again at generation time)."""
    _, g = load_golden("run_heat1d_imex_ie_step3A")  # tutorial/step_3/A_getting_statistics.py:43
    assert g["niter"].tolist() == [12] * 8
    _, g = load_golden("pfasst_step8A_heat1d")  # tutorial/step_8/A_visualize_residuals.py:56-58
    assert g["niter"].tolist() == [7] * 8 and float(g["err"]) < 6.1555e-05
    _, g = load_golden("run_heat1d_gi_lu_direct")
    assert g["niter"].tolist() == [10, 10, 10, 9, 9, 9, 8, 8, 8, 7, 7, 6, 6, 6, 5, 5, 5, 4, 4, 4]
    _, g = load_golden("run_heat3d_gi_minsrns_31")
    assert g["niter"].tolist() == [7, 7] and g["work_CG"].tolist() == [142, 140]
    assert float(g["uend_maxabs"]) == pytest.approx(1.5025598331, abs=1e-9)
    _, g = load_golden("run_allencahn_gi_lu_128")
    assert g["niter"].tolist() == [7, 7] and g["work_newton"].tolist() == [34, 36]
