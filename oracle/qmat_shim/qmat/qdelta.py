"""QDelta (SDC preconditioner) generators (stand-in for ``qmat.qdelta``).

Call sites: ``pySDC/core/sweeper.py:97-123`` (``QDELTA_GENERATORS[name](qGen=coll.generator, tLeft=...)``,
``genCoeffs(k=k)``, ``genCoeffs(k=k, dTau=True) -> (QDelta, dTau)``) and ``:262-276``
(``isKDependent()``, ``type(gen).__name__`` used again as a key of ``QDELTA_GENERATORS``).
"""
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


class MIN_SR_FLEX(QDeltaGenerator):
    def isKDependent(self):
        return True

    def computeQDelta(self, k=None):
        k = 1 if k is None else int(k)
        if k < 1:
            k = 1
        if k > self.M:
            raise NotImplementedError("MIN-SR-FLEX falls back to MIN-SR-S for k > M; not restated in the stand-in")
        return np.diag(self.nodes - self.tLeft) / k


# the reference looks generators up by alias AND by class name (sweeper.py:273,275)
MIN_SR_NS.__name__ = "MIN-SR-NS"
MIN_SR_FLEX.__name__ = "MIN-SR-FLEX"

QDELTA_GENERATORS = {
    "BE": BE, "IE": BE,
    "FE": FE, "EE": FE,
    "LU": LU,
    "PIC": PIC,
    "Exact": Exact, "EXACT": Exact,
    "BEPAR": BEPAR, "IEpar": BEPAR,
    "Jacobi": Jacobi, "Qpar": Jacobi,
    "MIN-SR-NS": MIN_SR_NS,
    "MIN-SR-FLEX": MIN_SR_FLEX,
}
