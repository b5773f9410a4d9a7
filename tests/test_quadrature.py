"""Collocation / QDelta coefficients of the product (pysdc_b200/quadrature.py) against the independently written qmat
stand-in of the oracle, plus the reference's own property tests (pySDC/tests/test_collocation.py:19-120,
tests/test_sweepers/test_preconditioners.py:15-207)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "qmat_shim"))

from pysdc_b200.quadrature import CollBase, make_qdelta_generator  # noqa: E402

QUADS = ["GAUSS", "RADAU-LEFT", "RADAU-RIGHT", "LOBATTO"]


@pytest.mark.parametrize("quad", QUADS)
@pytest.mark.parametrize("node_type", ["LEGENDRE", "EQUID"])
@pytest.mark.parametrize("M", [2, 3, 4, 5, 7, 8])
def test_against_oracle_shim(M, quad, node_type):
    from qmat import Q_GENERATORS

    c = CollBase(num_nodes=M, quad_type=quad, node_type=node_type)
    g = Q_GENERATORS["Collocation"](nNodes=M, nodeType=node_type, quadType=quad, tLeft=0, tRight=1)
    tol = 5e-15 if node_type == "LEGENDRE" else 1e-12  # equidistant interpolation is ill-conditioned for M = 8
    np.testing.assert_allclose(c.nodes, g.nodes, rtol=0, atol=2e-15)
    np.testing.assert_allclose(c.weights, g.weights, rtol=0, atol=tol)
    np.testing.assert_allclose(c.Qmat[1:, 1:], g.Q, rtol=0, atol=tol)
    assert c.order == g.order


@pytest.mark.parametrize("quad", QUADS)
@pytest.mark.parametrize("M", [2, 3, 5, 8])
def test_polynomial_exactness_and_S(M, quad):
    """test_collocation.py: weights integrate polynomials up to order-1 exactly; Q rows integrate to the nodes;
    Q = cumsum(S)."""
    for a, b in ((0.0, 1.0), (-0.3, 0.8)):
        c = CollBase(num_nodes=M, tleft=a, tright=b, quad_type=quad)
        for deg in range(c.order):
            p = np.polynomial.Polynomial(np.random.default_rng(deg).standard_normal(deg + 1))
            P = p.integ()
            assert abs(c.weights @ p(c.nodes) - (P(b) - P(a))) < 1e-13
        for deg in range(M):
            p = np.polynomial.Polynomial(np.random.default_rng(deg).standard_normal(deg + 1))
            P = p.integ()
            np.testing.assert_allclose(c.Qmat[1:, 1:] @ p(c.nodes), P(c.nodes) - P(a), rtol=0, atol=1e-13)
        np.testing.assert_allclose(np.cumsum(c.Smat[1:, 1:], axis=0), c.Qmat[1:, 1:], rtol=0, atol=1e-15)
        np.testing.assert_allclose(c.delta_m.sum(), c.nodes[-1] - a, atol=1e-15)


@pytest.mark.parametrize("name", ["IE", "EE", "LU", "PIC", "IEpar", "Qpar", "MIN-SR-NS", "MIN-SR-S", "MIN-SR-FLEX"])
@pytest.mark.parametrize("M", [2, 3, 4, 5])
def test_qdelta_against_oracle_shim(M, name):
    from qmat import Q_GENERATORS
    from qmat.qdelta import QDELTA_GENERATORS

    c = CollBase(num_nodes=M, quad_type="RADAU-RIGHT")
    g = Q_GENERATORS["Collocation"](nNodes=M, nodeType="LEGENDRE", quadType="RADAU-RIGHT", tLeft=0, tRight=1)
    mine, ref = make_qdelta_generator(name, c), QDELTA_GENERATORS[name](qGen=g, tLeft=0)
    assert mine.isKDependent() == ref.isKDependent()
    # k > M: MIN-SR-FLEX continues with the MIN-SR-S coefficients (a nonlinear solve: agreement to solver accuracy)
    for k in ([None] if not mine.isKDependent() else [1, 2, M, M + 1]):
        QD, dtau = ref.genCoeffs(k=k, dTau=True)
        np.testing.assert_allclose(mine.coeffs(k), QD, rtol=0, atol=1e-12 if name == "MIN-SR-S" or (k or 0) > M else 2e-15)
        np.testing.assert_allclose(mine.dtau(k), dtau, rtol=0, atol=2e-15)


@pytest.mark.parametrize("M", [2, 3, 4, 5])
def test_preconditioner_properties(M):
    """test_preconditioners.py: MIN-SR-NS diagonal and (I - QD^-1 Q) nilpotent-ish in the non-stiff limit (:15-41),
    LU nilpotency of (I - QD^-1 Q) in the stiff limit to 1e-14 (:134-154), FLEX product nilpotent (:48-73)."""
    c = CollBase(num_nodes=M, quad_type="RADAU-RIGHT")
    Q = c.Qmat[1:, 1:]
    I = np.eye(M)
    QD = make_qdelta_generator("LU", c).coeffs()
    assert np.allclose(np.tril(QD), QD)
    K = I - np.linalg.solve(QD, Q)  # stiff-limit iteration matrix
    assert np.max(np.abs(np.linalg.matrix_power(K, M))) < 1e-11
    QD = make_qdelta_generator("MIN-SR-NS", c).coeffs()
    assert np.allclose(np.diag(np.diag(QD)), QD)
    assert np.max(np.abs(np.linalg.eigvals(Q - QD))) < 1e-7 ** (1.0 / M) * 10  # non-stiff limit: spectral radius ~ 0
    # MIN-SR-S (test_preconditioners.py:15-41): diagonal, stiff-limit iteration matrix nilpotent
    QD = make_qdelta_generator("MIN-SR-S", c).coeffs()
    assert np.allclose(np.diag(np.diag(QD)), QD) and np.all(np.diag(QD) > 0)
    assert np.max(np.abs(np.linalg.eigvals(I - np.linalg.solve(QD, Q)))) < 2e-2 ** (1.0 if M > 4 else 2.0)
    flex = make_qdelta_generator("MIN-SR-FLEX", c)
    prod = I
    for k in range(1, M + 1):
        prod = (I - np.linalg.solve(flex.coeffs(k), Q)) @ prod
    assert np.max(np.abs(prod)) < 1e-9
