"""Collocation nodes, quadrature matrices and QDelta preconditioners (host-side set-up, tiny M x M matrices).

Mirrors what the reference obtains from the third-party ``qmat`` package through ``pySDC/core/collocation.py:48-108``
(``CollBase``) and ``pySDC/core/sweeper.py:97-123`` (``get_Qdelta_implicit`` / ``get_Qdelta_explicit``).  The
matrices never live on the GPU: they enter the collocation kernels as launch arguments.

Written independently of the test oracle's qmat stand-in (different algorithms: Jacobi-polynomial roots from
``scipy.special`` and exact integration in the Legendre basis instead of barycentric interpolation + Gauss
quadrature); ``tests/test_quadrature.py`` checks that the two agree to round-off and re-asserts the reference's
property tests (``pySDC/tests/test_collocation.py``, ``tests/test_sweepers/test_preconditioners.py``).
"""
import warnings

import numpy as np
import scipy.linalg
import scipy.special
from numpy.polynomial import legendre as _leg

from .errors import CollocationError, ParameterError

NODE_TYPES = ("LEGENDRE", "EQUID")
QUAD_TYPES = ("GAUSS", "RADAU-LEFT", "RADAU-RIGHT", "LOBATTO")


def _reference_nodes(M, node_type, quad_type):
    """Nodes on [-1, 1]."""
    if node_type == "EQUID":
        if quad_type == "GAUSS":
            return np.linspace(-1.0, 1.0, M + 2)[1:-1]
        if quad_type == "LOBATTO":
            return np.linspace(-1.0, 1.0, M)
        if quad_type == "RADAU-RIGHT":
            return np.linspace(-1.0, 1.0, M + 1)[1:]
        return np.linspace(-1.0, 1.0, M + 1)[:-1]
    # LEGENDRE: interior nodes are roots of Jacobi polynomials P^(a,b); included end points are appended exactly
    if quad_type == "GAUSS":
        return np.sort(scipy.special.roots_legendre(M)[0])
    if quad_type == "RADAU-RIGHT":
        inner = scipy.special.roots_jacobi(M - 1, 1.0, 0.0)[0] if M > 1 else np.array([])
        return np.concatenate((np.sort(inner), [1.0]))
    if quad_type == "RADAU-LEFT":
        inner = scipy.special.roots_jacobi(M - 1, 0.0, 1.0)[0] if M > 1 else np.array([])
        return np.concatenate(([-1.0], np.sort(inner)))
    if M < 2:
        raise CollocationError("LOBATTO needs at least two nodes")
    inner = scipy.special.roots_jacobi(M - 2, 1.0, 1.0)[0] if M > 2 else np.array([])
    return np.concatenate(([-1.0], np.sort(inner), [1.0]))


def _integration_matrix(x, upper):
    """Row i: integrals over [-1, upper[i]] of the Lagrange basis through the nodes x (all on [-1, 1]).

    The basis is expanded in Legendre polynomials (well-conditioned Vandermonde), which integrate in closed form:
    int_{-1}^{t} P_k = (P_{k+1}(t) - P_{k-1}(t)) / (2k+1),  int_{-1}^{t} P_0 = t + 1.
    """
    M = x.size
    V = _leg.legvander(x, M - 1)  # V[i, k] = P_k(x_i)
    t = np.asarray(upper, dtype=float)
    Pt = _leg.legvander(t, M)  # up to P_M
    I = np.empty((t.size, M))
    I[:, 0] = t + 1.0
    for k in range(1, M):
        I[:, k] = (Pt[:, k + 1] - Pt[:, k - 1]) / (2 * k + 1)
    # Lagrange basis coefficients C = V^{-1}  =>  integrals = I @ C
    return np.linalg.solve(V.T, I.T).T


class CollBase:
    """Same attributes as ``pySDC.core.collocation.CollBase`` (collocation.py:36-46)."""

    def __init__(self, num_nodes=None, tleft=0, tright=1, node_type="LEGENDRE", quad_type=None, **kwargs):
        if num_nodes is None or not num_nodes > 0:
            raise CollocationError("at least one quadrature node required, got %s" % num_nodes)
        if not tleft < tright:
            raise CollocationError("interval boundaries are corrupt, got %s and %s" % (tleft, tright))
        if node_type not in NODE_TYPES:
            raise CollocationError(f"node_type {node_type!r} not available (have {NODE_TYPES})")
        if quad_type not in QUAD_TYPES:
            raise CollocationError(f"quad_type {quad_type!r} not available (have {QUAD_TYPES})")
        M = int(num_nodes)
        self.num_nodes, self.tleft, self.tright = M, tleft, tright
        self.node_type, self.quad_type = node_type, quad_type
        self.left_is_node = quad_type in ("LOBATTO", "RADAU-LEFT")
        self.right_is_node = quad_type in ("LOBATTO", "RADAU-RIGHT")
        x = _reference_nodes(M, node_type, quad_type)
        half = 0.5 * (tright - tleft)
        self.nodes = half * (x + 1.0) + tleft
        if self.right_is_node:
            self.nodes[-1] = tright
        if self.left_is_node:
            self.nodes[0] = tleft
        self.weights = half * _integration_matrix(x, [1.0])[0]
        self.Qmat = np.zeros((M + 1, M + 1))
        self.Qmat[1:, 1:] = half * _integration_matrix(x, x)
        self.Smat = np.zeros((M + 1, M + 1))
        self.Smat[1, :] = self.Qmat[1, :]
        self.Smat[2:, :] = self.Qmat[2:, :] - self.Qmat[1:-1, :]
        self.delta_m = np.diff(np.concatenate(([tleft], self.nodes)))
        if node_type == "LEGENDRE":
            self.order = {"GAUSS": 2 * M, "RADAU-LEFT": 2 * M - 1, "RADAU-RIGHT": 2 * M - 1, "LOBATTO": 2 * M - 2}[quad_type]
        else:
            self.order = M


# ---------------------------------------------------------------------------------------------------------------------
# QDelta generators
# ---------------------------------------------------------------------------------------------------------------------
class QDeltaGenerator:
    """M x M coefficients of a preconditioner; ``coeffs(k)`` may depend on the sweep index k."""

    k_dependent = False

    def __init__(self, coll):
        self.coll = coll
        self.M = coll.num_nodes
        self.Q = coll.Qmat[1:, 1:]
        self.nodes = coll.nodes
        self.tleft = coll.tleft

    def isKDependent(self):
        return self.k_dependent

    def coeffs(self, k=None):
        raise NotImplementedError

    def dtau(self, k=None):
        return np.zeros(self.M)


class _IE(QDeltaGenerator):
    def coeffs(self, k=None):
        return np.tril(np.tile(self.coll.delta_m, (self.M, 1)))


class _EE(QDeltaGenerator):
    def coeffs(self, k=None):
        QD = np.zeros((self.M, self.M))
        for i in range(1, self.M):
            QD[i, :i] = self.coll.delta_m[1 : i + 1]
        return QD

    def dtau(self, k=None):
        return np.full(self.M, self.coll.delta_m[0])


class _LU(QDeltaGenerator):
    def coeffs(self, k=None):
        return scipy.linalg.lu(self.Q.T)[2].T


class _PIC(QDeltaGenerator):
    def coeffs(self, k=None):
        return np.zeros((self.M, self.M))


class _IEpar(QDeltaGenerator):
    def coeffs(self, k=None):
        return np.diag(self.nodes - self.tleft)


class _Qpar(QDeltaGenerator):
    def coeffs(self, k=None):
        return np.diag(np.diag(self.Q))


class _MinSrNs(QDeltaGenerator):
    def coeffs(self, k=None):
        return np.diag(self.nodes - self.tleft) / self.M


def _min_sr_s_diagonal(M, node_type, quad_type):
    """Diagonal d (on [0, 1]) making the stiff-limit iteration matrix K = I - diag(d)^-1 Q nilpotent, i.e. all
    coefficients of det(lambda I - K) but the leading one vanish.  Equivalent published form (qmat ``MIN-SR-S``):
    det((1 - z) I + z diag(1/d) Q) = 1 at M distinct z (the nodes).  The root near the power law a * nodes**b / M is
    taken, found incrementally in M as the published algorithm does, so that the same branch is selected."""
    import scipy.optimize

    first_is_zero = quad_type in ("LOBATTO", "RADAU-LEFT")
    a = b = None
    d = None
    for m in range(2 if first_is_zero else 1, M + 1):
        c = CollBase(num_nodes=m, tleft=0.0, tright=1.0, node_type=node_type, quad_type=quad_type)
        Q, t = c.Qmat[1:, 1:], c.nodes
        if first_is_zero:
            Q, t = Q[1:, 1:], t[1:]
        k = t.size
        if k == 1:
            d = np.array([Q[0, 0]])
        else:
            eye = np.eye(k)

            def conditions(x):
                G = Q / np.asarray(x)[:, None]  # diag(1/x) Q
                return np.array([np.linalg.det((1.0 - z) * eye + z * G) - 1.0 for z in t])

            start = t / m if a is None else a * t**b / m
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)  # 'xtol too small': converged to round-off
                d = scipy.optimize.fsolve(conditions, start, xtol=1e-15)
        if m > 1:
            target = d * m
            a, b = scipy.optimize.minimize(lambda ab: np.linalg.norm(ab[0] * t**ab[1] - target), [1.0, 1.0],
                                           method="nelder-mead").x
    return np.concatenate(([0.0], d)) if first_is_zero else d


class _MinSrS(QDeltaGenerator):
    """MIN-SR-S: diagonal, minimal spectral radius of the sweep's iteration matrix in the stiff limit."""

    def coeffs(self, k=None):
        d = _min_sr_s_diagonal(self.M, self.coll.node_type, self.coll.quad_type)
        return np.diag(d) * (self.coll.tright - self.coll.tleft)


class _MinSrFlex(_MinSrS):
    k_dependent = True

    def coeffs(self, k=None):
        k = 1 if k is None or k < 1 else int(k)
        if k > self.M:  # qmat: beyond M sweeps MIN-SR-FLEX continues with the MIN-SR-S coefficients
            return super().coeffs()
        return np.diag(self.nodes - self.tleft) / k


QDELTA_GENERATORS = {
    "IE": _IE, "BE": _IE,
    "EE": _EE, "FE": _EE,
    "LU": _LU,
    "PIC": _PIC,
    "IEpar": _IEpar, "BEPAR": _IEpar,
    "Qpar": _Qpar, "Jacobi": _Qpar,
    "MIN-SR-NS": _MinSrNs,
    "MIN-SR-S": _MinSrS,
    "MIN-SR-FLEX": _MinSrFlex,
}


def make_qdelta_generator(name, coll):
    try:
        return QDELTA_GENERATORS[name](coll)
    except KeyError:
        raise ParameterError(f"QDelta type {name!r} not available (have {sorted(QDELTA_GENERATORS)})") from None
