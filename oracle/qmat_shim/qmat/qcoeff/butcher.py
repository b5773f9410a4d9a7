"""Only present so that ``pySDC/implementations/sweeper_classes/Runge_Kutta.py`` can be imported (the reference's
adaptivity / embedded-error convergence controllers import it; the module builds every scheme's class attributes at
import time, :497-816).  The Runge-Kutta sweepers are outside the SDC sweep path and their Butcher tables are not
restated: every scheme is a one-stage placeholder filled with NaN, so a run that actually used one would be visibly
wrong instead of silently different."""
import numpy as np


class _Placeholder:
    nodes = np.array([np.nan])
    weights = np.array([np.nan])
    Q = np.array([[np.nan]])

    def genCoeffs(self, embedded=False):
        weights = np.array([[np.nan], [np.nan]]) if embedded else self.weights
        return self.nodes, weights, self.Q


class _Placeholders(dict):
    def __missing__(self, key):
        return _Placeholder


RK_SCHEMES = _Placeholders()
