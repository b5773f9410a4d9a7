"""Barycentric Lagrange interpolation / integration (stand-in for ``qmat.lagrange``).

Call site: ``pySDC/core/base_transfer.py:90-91`` (``LagrangeApproximation(c_nodes).getInterpolationMatrix(f_nodes)``)
and, inside this stand-in, the Q / weights of the collocation generator.
"""
import numpy as np


class LagrangeApproximation:
    def __init__(self, points):
        x = np.asarray(points, dtype=float).ravel()
        self.points = x
        d = x[:, None] - x[None, :]
        np.fill_diagonal(d, 1.0)
        # barycentric weights w_j = 1 / prod_{k != j} (x_j - x_k); product order fixed (k ascending)
        self.weights = 1.0 / np.prod(d, axis=1)

    @property
    def n(self):
        return self.points.size

    def getInterpolationMatrix(self, times):
        t = np.asarray(times, dtype=float).ravel()
        x, w = self.points, self.weights
        diff = t[:, None] - x[None, :]
        hit = diff == 0.0
        diff[hit] = 1.0
        P = w[None, :] / diff
        P /= P.sum(axis=1)[:, None]
        rows = hit.any(axis=1)
        P[rows] = hit[rows].astype(float)
        return P

    def getIntegrationMatrix(self, intervals):
        """Row i = integrals of every Lagrange basis polynomial over intervals[i] = (a, b);
        Gauss-Legendre with ceil(n/2)+1 points is exact for the degree n-1 basis."""
        n = self.n
        gx, gw = np.polynomial.legendre.leggauss(n // 2 + 2)
        out = np.zeros((len(intervals), n))
        for i, (a, b) in enumerate(intervals):
            if b == a:
                continue
            t = 0.5 * (b - a) * (gx + 1.0) + a
            out[i] = 0.5 * (b - a) * (gw @ self.getInterpolationMatrix(t))
        return out
