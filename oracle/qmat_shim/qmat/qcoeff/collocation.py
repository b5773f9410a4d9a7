"""Collocation nodes / weights / Q matrix (stand-in for ``qmat.qcoeff.collocation.Collocation``).

Call site: ``pySDC/core/collocation.py:73-75`` (``nNodes, nodeType, quadType, tLeft, tRight`` →
``.nodes .weights .Q .order``).  Algorithm (published in the qmat docs): nodes are roots of
(combinations of) Legendre polynomials on [-1, 1] mapped to [tLeft, tRight]; weights and Q are exact
integrals of the Lagrange basis through those nodes.
"""
import numpy as np
from numpy.polynomial import legendre as L

from . import QGenerator
from ..lagrange import LagrangeApproximation


def _legendre_nodes(M, quadType):
    """Nodes on [-1, 1]."""
    e = lambda n: np.eye(n + 1)[n]  # coefficient vector of P_n in the Legendre basis
    if quadType == "GAUSS":
        x = L.legroots(e(M))
    elif quadType == "RADAU-RIGHT":
        if M == 1:
            return np.array([1.0])
        c = np.zeros(M + 1)
        c[M - 1], c[M] = 1.0, -1.0  # P_{M-1} - P_M vanishes at +1
        x = L.legroots(c)
    elif quadType == "RADAU-LEFT":
        if M == 1:
            return np.array([-1.0])
        c = np.zeros(M + 1)
        c[M - 1], c[M] = 1.0, 1.0  # P_{M-1} + P_M vanishes at -1
        x = L.legroots(c)
    elif quadType == "LOBATTO":
        if M < 2:
            raise ValueError("LOBATTO needs at least two nodes")
        inner = L.legroots(L.legder(e(M - 1))) if M > 2 else np.array([])
        x = np.concatenate(([-1.0], inner, [1.0]))
    else:
        raise ValueError(f"unknown quadType {quadType!r}")
    x = np.sort(np.real(x))
    # pin the included end points exactly
    if quadType in ("RADAU-RIGHT", "LOBATTO"):
        x[-1] = 1.0
    if quadType in ("RADAU-LEFT", "LOBATTO"):
        x[0] = -1.0
    return x


def _equid_nodes(M, quadType):
    if quadType == "GAUSS":
        return np.linspace(-1, 1, M + 2)[1:-1]
    if quadType == "LOBATTO":
        return np.linspace(-1, 1, M)
    if quadType == "RADAU-RIGHT":
        return np.linspace(-1, 1, M + 1)[1:]
    if quadType == "RADAU-LEFT":
        return np.linspace(-1, 1, M + 1)[:-1]
    raise ValueError(f"unknown quadType {quadType!r}")


class Collocation(QGenerator):
    def __init__(self, nNodes=4, nodeType="LEGENDRE", quadType="RADAU-RIGHT", tLeft=0, tRight=1, **_):
        if quadType is None:
            raise ValueError("quadType must be given")
        if nodeType == "LEGENDRE":
            x = _legendre_nodes(nNodes, quadType)
        elif nodeType == "EQUID":
            x = _equid_nodes(nNodes, quadType)
        else:
            raise NotImplementedError(f"nodeType {nodeType!r} is not covered by the qmat stand-in")
        self.nodeType, self.quadType = nodeType, quadType
        self.tLeft, self.tRight = tLeft, tRight
        a, b = float(tLeft), float(tRight)
        self.nodes = 0.5 * (b - a) * (x + 1.0) + a
        approx = LagrangeApproximation(self.nodes)
        self.weights = approx.getIntegrationMatrix([(a, b)]).ravel()
        self.Q = approx.getIntegrationMatrix([(a, tau) for tau in self.nodes])

    @property
    def order(self):
        M = self.nodes.size
        if self.nodeType != "LEGENDRE":
            return M
        return {"GAUSS": 2 * M, "RADAU-LEFT": 2 * M - 1, "RADAU-RIGHT": 2 * M - 1, "LOBATTO": 2 * M - 2}[self.quadType]
