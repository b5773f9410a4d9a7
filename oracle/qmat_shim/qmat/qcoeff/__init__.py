"""Q-coefficient generators (collocation only; see package docstring)."""
import numpy as np


class QGenerator:
    """Base class: the reference reads ``S`` from *this* class on purpose
    (``super(self.generator.__class__, self.generator).S``, ``pySDC/core/collocation.py:100``):
    node-to-node weights as row differences of Q."""

    @property
    def nNodes(self):
        return np.size(self.nodes)

    @property
    def S(self):
        Q = np.asarray(self.Q, dtype=float)
        S = Q.copy()
        S[1:] = Q[1:] - Q[:-1]
        return S
