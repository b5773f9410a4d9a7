"""QDelta (SDC preconditioner) generators (stand-in for ``qmat.qdelta``).

Call sites: ``pySDC/core/sweeper.py:97-123`` (``QDELTA_GENERATORS[name](qGen=coll.generator, tLeft=...)``,
``genCoeffs(k=k)``, ``genCoeffs(k=k, dTau=True) -> (QDelta, dTau)``) and ``:262-276``
(``isKDependent()``, ``type(gen).__name__`` used again as a key of ``QDELTA_GENERATORS``).
"""
import warnings

import numpy as np
import scipy.linalg as spl


class QDeltaGenerator:
    def __init__(self, qGen=None, tLeft=0.0, Q=None, nodes=None, **_):
        if qGen is not None:
            Q, nodes = qGen.Q, qGen.nodes
        self.Q = np.asarray(Q, dtype=float)
        self.nodes = np.asarray(nodes, dtype=float)
        self.tLeft = float(tLeft)

    @property
    def M(self):
        return self.nodes.size

    def isKDependent(self):
        return False

    def computeQDelta(self, k=None):
        raise NotImplementedError

    def computeDTau(self, k=None):
        return np.zeros(self.M)

    def genCoeffs(self, k=None, dTau=False):
        QD = np.array(self.computeQDelta(k), dtype=float)
        if dTau:
            return QD, np.array(self.computeDTau(k), dtype=float)
        return QD

    def _deltas(self):
        return np.diff(np.concatenate(([self.tLeft], self.nodes)))


class BE(QDeltaGenerator):  # implicit Euler between nodes: row i holds the increments up to node i
    def computeQDelta(self, k=None):
        d = self._deltas()
        return np.tril(np.tile(d, (self.M, 1)))


class FE(QDeltaGenerator):  # explicit Euler: strictly lower, first increment goes to the tLeft column
    def computeQDelta(self, k=None):
        d = self._deltas()
        QD = np.zeros((self.M, self.M))
        for i in range(1, self.M):
            QD[i, :i] = d[1 : i + 1]
        return QD

    def computeDTau(self, k=None):
        return np.full(self.M, self._deltas()[0])


class LU(QDeltaGenerator):  # U^T of the (pivoted, P ignored) LU factorisation of Q^T
    def computeQDelta(self, k=None):
        _, _, U = spl.lu(self.Q.T)
        return U.T


class PIC(QDeltaGenerator):
    def computeQDelta(self, k=None):
        return np.zeros((self.M, self.M))


class Exact(QDeltaGenerator):
    def computeQDelta(self, k=None):
        return self.Q.copy()


class BEPAR(QDeltaGenerator):
    def computeQDelta(self, k=None):
        return np.diag(self.nodes - self.tLeft)


class Jacobi(QDeltaGenerator):
    def computeQDelta(self, k=None):
        return np.diag(np.diag(self.Q))


class MIN_SR_NS(QDeltaGenerator):
    def computeQDelta(self, k=None):
        return np.diag(self.nodes - self.tLeft) / self.M


class MIN_SR_S(QDeltaGenerator):
    """Diagonal QDelta with minimal spectral radius of the stiff-limit iteration matrix I - QDelta^-1 Q (published
    algorithm of qmat / pySDC <= 5.4 ``get_Qdelta_implicit('MIN-SR-S')``): the coefficients d solve the nilpotency
    conditions  det((1 - z) I + z diag(1/d) Q) = 1  at z = the nodes; the nonlinear solve is started from a power law
    a * nodes**b / M fitted incrementally to the solutions for fewer nodes."""

    def __init__(self, qGen=None, tLeft=0.0, Q=None, nodes=None, nodeType="LEGENDRE", quadType="RADAU-RIGHT", **kw):
        super().__init__(qGen=qGen, tLeft=tLeft, Q=Q, nodes=nodes, **kw)
        self.nodeType = getattr(qGen, "nodeType", nodeType)
        self.quadType = getattr(qGen, "quadType", quadType)

    def _coeffs_for(self, M, a, b):
        from scipy import optimize

        from .qcoeff.collocation import Collocation

        coll = Collocation(nNodes=M, nodeType=self.nodeType, quadType=self.quadType, tLeft=0.0, tRight=1.0)
        QM, nodesM = coll.Q, coll.nodes
        first_is_zero = self.quadType in ("LOBATTO", "RADAU-LEFT")
        if first_is_zero:
            QM, nodesM = QM[1:, 1:], nodesM[1:]
        nC = nodesM.size
        if nC == 1:
            coeffs = np.diag(QM).copy()
        else:
            def nilpotency(c):
                return np.array([np.linalg.det((1 - z) * np.eye(nC) + z * np.diag(1 / np.asarray(c)) @ QM) - 1
                                 for z in nodesM])

            c0 = nodesM / M if a is None else a * nodesM**b / M
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)  # 'xtol too small': converged to round-off
                coeffs = optimize.fsolve(nilpotency, c0, xtol=1e-15)
        if first_is_zero:
            coeffs, nodesM = np.concatenate(([0.0], coeffs)), np.concatenate(([0.0], nodesM))
        return coeffs, nodesM

    def computeQDelta(self, k=None):
        from scipy import optimize

        a = b = None
        m0 = 2 if self.quadType in ("LOBATTO", "RADAU-LEFT") else 1
        coeffs = None
        for m in range(m0, self.M + 1):
            coeffs, nodes = self._coeffs_for(m, a, b)
            if m > 1:
                target = coeffs * m
                a, b = optimize.minimize(lambda ab: np.linalg.norm(ab[0] * nodes**ab[1] - target), [1.0, 1.0],
                                         method="nelder-mead").x
        # coefficients are computed on [0, 1]; scale to the step the generator was built for
        scale = (self.nodes[-1] - self.tLeft) / 1.0 if self.quadType in ("LOBATTO", "RADAU-RIGHT") else 1.0
        return np.diag(coeffs) * scale


class MIN_SR_FLEX(MIN_SR_S):
    def isKDependent(self):
        return True

    def computeQDelta(self, k=None):
        k = 1 if k is None else int(k)
        if k < 1:
            k = 1
        if k > self.M:  # beyond M sweeps the flexible scheme continues with the stiff-limit coefficients
            return super().computeQDelta()
        return np.diag(self.nodes - self.tLeft) / k


# the reference looks generators up by alias AND by class name (sweeper.py:273,275)
MIN_SR_NS.__name__ = "MIN-SR-NS"
MIN_SR_FLEX.__name__ = "MIN-SR-FLEX"
MIN_SR_S.__name__ = "MIN-SR-S"

QDELTA_GENERATORS = {
    "BE": BE, "IE": BE,
    "FE": FE, "EE": FE,
    "LU": LU,
    "PIC": PIC,
    "Exact": Exact, "EXACT": Exact,
    "BEPAR": BEPAR, "IEpar": BEPAR,
    "Jacobi": Jacobi, "Qpar": Jacobi,
    "MIN-SR-NS": MIN_SR_NS,
    "MIN-SR-S": MIN_SR_S,
    "MIN-SR-FLEX": MIN_SR_FLEX,
}
