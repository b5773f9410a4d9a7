"""Space-time transfer between the levels of a step (MLSDC / PFASST), on device fields.

* ``mesh_to_mesh`` — same name, constructor and parameters (``periodic, equidist_nested, iorder, rorder``) as
  ``pySDC/implementations/transfer_classes/TransferMesh.py:9-218``.  The reference assembles sparse Kronecker products
  of 1-D barycentric interpolation matrices (``helpers/transfer_helper.py:139-247``) and multiplies flattened fields
  with them; here the same 1-D operators are built on the host (small: ``n_f x k`` entries), uploaded once in ELL form
  and applied axis by axis by the ``sdcb200_axis_apply`` kernel (``kron(A, B) vec(G) = vec(A G B^T)``).
* ``BaseTransfer`` — ``pySDC/core/base_transfer.py:30-251``: restriction with FAS correction ``tau``, prolongation of the
  coarse correction (``prolong`` / ``prolong_f``), node-to-node transfer matrices ``Rcoll`` / ``Pcoll``.  Written against
  the datatype operators, so it is also what pySDC's own ``BaseTransfer`` does when it drives these classes.
"""
import numpy as np

from .backend import get_backend
from .errors import ParameterError, TransferError, UnlockError


def _lagrange_weights(nodes, p):
    """Values at ``p`` of the Lagrange basis polynomials on ``nodes`` (what ``BarycentricInterpolator`` evaluates in
    transfer_helper.py:224-229)."""
    nodes = np.asarray(nodes, dtype=float)
    w = np.ones(nodes.size)
    for j in range(nodes.size):
        for m in range(nodes.size):
            if m != j:
                w[j] *= (p - nodes[m]) / (nodes[j] - nodes[m])
    return w


def _border_padding(grid, l, r):
    """Mirror padding of a grid (transfer_helper.py:250-273)."""
    out = np.zeros(grid.size + l + r)
    for i in range(l):
        out[i] = 2 * grid[0] - grid[l - i]
    for j in range(r):
        out[-j - 1] = 2 * grid[-1] - grid[-r + j - 1]
    out[l: l + grid.size] = grid
    return out


def interpolation_matrix_1d(fine_grid, coarse_grid, k=2, periodic=False, pad=1, equidist_nested=True):
    """Dense ``n_f x n_c`` interpolation matrix for equidistant nested grids (transfer_helper.py:139-247, the
    ``equidist_nested`` branches, which is what ``mesh_to_mesh`` uses by default)."""
    if not equidist_nested:
        raise TransferError("only equidistant nested grids are implemented (equidist_nested=True)")
    n_f, n_c = fine_grid.size, coarse_grid.size
    if periodic:
        M = np.zeros((n_f, n_c))
        for i, p in enumerate(fine_grid):
            if i % 2 == 0:
                M[i, i // 2] = 1.0
                continue
            cpos, offset = i // 2, k // 2
            nn = sorted((cpos - offset + 1 + j) % n_c for j in range(k))
            nodes = coarse_grid[nn].copy()
            d = np.diff(nn)
            if not np.all(d == 1):  # the neighbour set wraps around: continue the grid periodically
                shift, cont = 0.0, [coarse_grid[nn[0]]]
                for idx, dd in zip(nn[1:], d):
                    if dd != 1:
                        shift = -1.0
                    cont.append(coarse_grid[idx] + shift)
                nodes = np.asarray(cont)
            if p > np.mean(fine_grid) and not (nodes[0] <= p <= nodes[-1]):
                nodes = nodes + 1.0
            M[i, nn] = _lagrange_weights(nodes, p)
        return M
    M = np.zeros((n_f, n_c + 2 * pad))
    padded = _border_padding(coarse_grid, pad, pad)
    for i, p in enumerate(fine_grid):
        if i % 2 != 0:
            M[i, (i - 1) // 2 + 1] = 1.0
            continue
        cpos, offset = i // 2, k // 2
        nn = []
        for j in range(k):
            v = cpos - offset + 1 + j
            if v < 0:
                v += k
            elif v > n_c + 1:
                v -= k
            nn.append(v)
        nn = sorted(nn)
        M[i, nn] = _lagrange_weights(padded[nn], p)
    return M[:, pad:-pad] if pad > 0 else M


def to_ell(M):
    """Dense operator -> (weights[n_out, width], columns[n_out, width]) with -1 for unused slots."""
    M = np.asarray(M)
    nz = [np.flatnonzero(row) for row in M]
    width = max(1, max(len(c) for c in nz))
    W = np.zeros((M.shape[0], width))
    C = -np.ones((M.shape[0], width), dtype=np.int32)
    for i, c in enumerate(nz):
        W[i, : len(c)] = M[i, c]
        C[i, : len(c)] = c
    return W, C


class mesh_to_mesh:
    """Restriction / prolongation between nested FD grids with dirichlet-zero or periodic boundaries."""

    def __init__(self, fine_prob, coarse_prob, params):
        self.params = type("Pars", (), dict(periodic=False, equidist_nested=True, iorder=2, rorder=2))()
        for k, v in params.items():
            if not hasattr(self.params, k):
                raise ParameterError(f"unknown space transfer parameter {k!r}")
            setattr(self.params, k, v)
        self.fine_prob, self.coarse_prob = fine_prob, coarse_prob
        if self.params.rorder % 2 != 0:
            raise TransferError("Need even order for restriction")
        if self.params.iorder % 2 != 0:
            raise TransferError("Need even order for interpolation")
        nf, nc = fine_prob.nvars, coarse_prob.nvars
        nf = (nf,) if isinstance(nf, int) else tuple(nf)
        nc = (nc,) if isinstance(nc, int) else tuple(nc)
        if len(nf) != len(nc):
            raise TransferError("nvars parameter of fine and coarse level needs to have the same length")
        self._nf, self._nc = nf, nc
        self._be = get_backend()
        self.identity = nf == nc
        if self.identity:
            return
        per = self.params.periodic
        off = 0 if per else 1
        fine_grid = np.array([(i + off) * fine_prob.dx for i in range(nf[0])])
        coarse_grid = np.array([(i + off) * coarse_prob.dx for i in range(nc[0])])
        P1 = interpolation_matrix_1d(fine_grid, coarse_grid, k=self.params.iorder, periodic=per,
                                     equidist_nested=self.params.equidist_nested)
        factor = 0.5 if self.params.rorder > 0 else 1.0
        if self.params.iorder == self.params.rorder:
            R1 = factor * P1.T
        else:
            R1 = factor * interpolation_matrix_1d(fine_grid, coarse_grid, k=self.params.rorder, periodic=per,
                                                  equidist_nested=self.params.equidist_nested).T
        self.Pspace_1d, self.Rspace_1d = P1, R1  # dense 1-D operators (host); N-D operators are their Kronecker powers
        self._P = self._be.upload_operator(*to_ell(P1))
        self._R = self._be.upload_operator(*to_ell(R1))
        self._tmp = {}

    # ---- one single-component field through the 1-D operator along every axis ------------------------------------
    def _apply(self, op, src, dst):
        """dst <- (op x op x ...) src, axis by axis starting with the contiguous one (x)."""
        ls, ld = src.layout, dst.layout
        ndim, n_in, n_out, Ps, Pd = ls.ndim, ls.n, ld.n, ls.P, ld.P
        be = self._be
        if ndim == 1:
            be.axis_apply(op, 1, n_out, 1, src.flat, 0, 1, dst.flat, 0, 1)
            return
        key = (id(op), ndim)
        if key not in self._tmp:
            # intermediates: x done (source rows x n_out columns), then (3-D) x and y done
            rows = n_in if ndim == 2 else n_in * Ps
            self._tmp[key] = [be.zeros(rows * Pd), be.zeros(n_in * Pd * Pd) if ndim == 3 else None]
        t1, t2 = self._tmp[key]
        if ndim == 2:
            be.axis_apply(op, n_in, n_out, 1, src.flat, Ps, 1, t1, Pd, 1)            # along x, per row
            be.axis_apply(op, 1, n_out, n_out, t1, 0, Pd, dst.flat, 0, Pd)          # along y
            return
        # 3-D: along x for every row of every owned plane (wall rows included: they are zero and stay zero),
        # along y per plane, along z
        be.axis_apply(op, n_in * Ps, n_out, 1, src.flat, Ps, 1, t1, Pd, 1)
        be.axis_apply(op, n_in, n_out, n_out, t1, Ps * Pd, Pd, t2, Pd * Pd, Pd)
        be.axis_apply(op, 1, n_out, Pd * Pd, t2, 0, Pd * Pd, dst.flat, 0, Pd * Pd)

    def _transfer(self, X, target_init, op):
        Y = type(X)(target_init)
        if self.identity:
            Y[:] = X
            return Y
        comps = type(X).components
        if comps:
            for c in comps:
                self._apply(op, getattr(X, c), getattr(Y, c))
        else:
            self._apply(op, X, Y)
        return Y

    def restrict(self, F):
        """TransferMesh.py:149-183."""
        return self._transfer(F, self.coarse_prob.init, None if self.identity else self._R)

    def prolong(self, G):
        """TransferMesh.py:185-218."""
        return self._transfer(G, self.fine_prob.init, None if self.identity else self._P)


class BaseTransfer:
    """core/base_transfer.py:30-251."""

    def __init__(self, fine_level, coarse_level, base_transfer_params, space_transfer_class, space_transfer_params):
        self.params = type("Pars", (), dict(finter=False))()
        for k, v in base_transfer_params.items():
            setattr(self.params, k, v)
        self.fine, self.coarse = fine_level, coarse_level
        fn, cn = self.fine.sweep.coll.nodes, self.coarse.sweep.coll.nodes
        if len(fn) == len(cn):
            self.Pcoll = np.eye(len(fn))
            self.Rcoll = np.eye(len(fn))
        else:
            self.Pcoll = self.get_transfer_matrix_Q(fn, cn)
            self.Rcoll = self.get_transfer_matrix_Q(cn, fn)
        self.space_transfer = space_transfer_class(fine_prob=self.fine.prob, coarse_prob=self.coarse.prob,
                                                   params=space_transfer_params)

    @staticmethod
    def get_transfer_matrix_Q(f_nodes, c_nodes):
        """Lagrange interpolation matrix from ``c_nodes`` to ``f_nodes`` (base_transfer.py:79-91)."""
        return np.array([_lagrange_weights(c_nodes, p) for p in f_nodes])

    # ---- node-to-node combinations as fused launches -----------------------------------------------------------------
    @staticmethod
    def _parts(x):
        comps = type(x).components
        return [getattr(x, c) for c in comps] if comps else [x]

    def _combine(self, weights, fields, like_init, dtype, base=None, subtract=None):
        """out[n] = (base[n] if given) + sum_m weights[n, m] * fields[m] (- subtract[n] if given) for all n at once:
        one fused collocation launch per component instead of the reference's nested axpy loops."""
        be = get_backend()
        weights = np.atleast_2d(np.asarray(weights, dtype=float))
        outs = [dtype(like_init) for _ in range(weights.shape[0])]
        W = weights if subtract is None else np.hstack([weights, -np.eye(weights.shape[0])])
        for c in range(len(self._parts(outs[0]))):
            ins = [self._parts(f)[c].flat for f in fields]
            if subtract is not None:
                ins += [self._parts(t)[c].flat for t in subtract]
            adds = None if base is None else [self._parts(b)[c].flat for b in base]
            be.colloc_apply(W, ins, None, adds, [self._parts(o)[c].flat for o in outs])
        return outs

    def restrict(self):
        """Space-time restriction with FAS correction (base_transfer.py:93-168): coarse values = Rcoll x (space-
        restricted fine values), coarse right-hand sides re-evaluated, tau = R(Q_f F_f) - Q_c F_c (+ restricted fine
        tau)."""
        F, G = self.fine, self.coarse
        PG, SF, SG = G.prob, F.sweep, G.sweep
        if not F.status.unlocked:
            raise UnlockError("fine level is still locked, cannot use data from there")
        MF, MG = SF.coll.num_nodes, SG.coll.num_nodes
        R = self.space_transfer.restrict
        G.u[0] = R(F.u[0])
        G.u[1:] = self._combine(self.Rcoll, [R(F.u[m]) for m in range(1, MF + 1)], PG.init, PG.dtype_u)
        times = [G.time] + [G.time + G.dt * SG.coll.nodes[m] for m in range(MG)]
        for m in range(MG + 1):
            G.f[m] = PG.eval_f(G.u[m], times[m])
        tau_coarse = G.sweep.integrate()
        restricted = [R(t) for t in F.sweep.integrate()]
        G.tau[:] = self._combine(self.Rcoll, restricted, PG.init, PG.dtype_u, subtract=tau_coarse)
        if F.tau[0] is not None:
            G.tau[:] = self._combine(self.Rcoll, [R(t) for t in F.tau], PG.init, PG.dtype_u, base=G.tau)
        for m in range(1, MG + 1):
            G.uold[m] = PG.dtype_u(G.u[m])
            G.fold[m] = PG.dtype_f(G.f[m])
        G.status.unlocked = True

    def _correct(self, fine_list, coarse_new, coarse_old, init, dtype):
        """fine[n] += sum_m Pcoll[n, m] * prolong(coarse_new[m] - coarse_old[m])."""
        MG = self.coarse.sweep.coll.num_nodes
        delta = [self.space_transfer.prolong(coarse_new[m] - coarse_old[m]) for m in range(1, MG + 1)]
        fine_list[1:] = self._combine(self.Pcoll, delta, init, dtype, base=fine_list[1:])

    def prolong(self):
        """Coarse correction of the fine values, fine right-hand sides re-evaluated (base_transfer.py:170-215)."""
        F, G = self.fine, self.coarse
        if not G.status.unlocked:
            raise UnlockError("coarse level is still locked, cannot use data from there")
        PF, SF = F.prob, F.sweep
        self._correct(F.u, G.u, G.uold, PF.init, PF.dtype_u)
        for m in range(1, SF.coll.num_nodes + 1):
            F.f[m] = PF.eval_f(F.u[m], F.time + F.dt * SF.coll.nodes[m - 1])

    def prolong_f(self):
        """Coarse correction of values AND right-hand sides, no re-evaluation (base_transfer.py:217-251)."""
        F, G = self.fine, self.coarse
        if not G.status.unlocked:
            raise UnlockError("coarse level is still locked, cannot use data from there")
        PF = F.prob
        self._correct(F.u, G.u, G.uold, PF.init, PF.dtype_u)
        self._correct(F.f, G.f, G.fold, PF.init, PF.dtype_f)
