"""TEST-ONLY numpy stand-in for ``pysdc_b200.backend.CudaBackend``.

It lets the CPU test-suite (no GPU in the build container) drive the *host* logic of the package — datatypes, problem
classes, sweepers, controller, the pySDC plug-in — end to end against the golden fixtures.  It is never imported by the
package; the product backend raises ``BackendError`` when the CUDA library or a GPU is missing.  Each method mirrors the
semantics of the C entry point of the same name in ``include/sdc_b200.h`` (plain numpy, independent of ``oracle/``).
"""
import numpy as np
import torch


def _np(t):
    return t.numpy()  # CPU tensors share memory with numpy


class NumpyBackend:
    name = "numpy-test-double"

    def __init__(self):
        self.device = torch.device("cpu")
        self.launches = 0

    def zeros(self, count, dtype=torch.float64):
        return torch.zeros(count, dtype=dtype)

    def synchronize(self):
        pass

    def device_info(self):
        return dict(sm_count=0, cc=(0, 0), solver_ctas=0)

    # ---- helpers ----------------------------------------------------------------------------------------------------
    @staticmethod
    def _grid(lay, t):
        """Writable numpy view of the grid points of a flat volume tensor."""
        return _np(lay.interior(t))

    @staticmethod
    def _lap_sum(x, periodic):
        """Sum of the 2*ndim neighbours (zero outside for Dirichlet)."""
        out = np.zeros_like(x)
        for ax in range(x.ndim):
            if periodic:
                out += np.roll(x, 1, axis=ax) + np.roll(x, -1, axis=ax)
            else:
                lo = [slice(None)] * x.ndim
                hi = [slice(None)] * x.ndim
                lo[ax], hi[ax] = slice(1, None), slice(None, -1)
                out[tuple(lo)] += x[tuple(hi)]
                out[tuple(hi)] += x[tuple(lo)]
        return out

    # ---- K6 ---------------------------------------------------------------------------------------------------------
    def maxabs_async(self, x, out):
        self.launches += 1
        out[0] = float(np.max(np.abs(_np(x)))) if x.numel() else 0.0

    def maxabs(self, x):
        out = torch.zeros(1, dtype=torch.float64)
        self.maxabs_async(x, out)
        return float(out[0])

    def axpby(self, a, x, b, y, out):
        self.launches += 1
        r = a * _np(x)
        if y is not None:
            r = r + b * _np(y)
        _np(out)[:] = r

    # ---- K1 ---------------------------------------------------------------------------------------------------------
    def colloc_apply(self, W, ins, base, adds, outs):
        self.launches += 1
        W = np.asarray(W, dtype=float).reshape(len(outs), len(ins))
        vals = [_np(t).copy() for t in ins]
        b = None if base is None else _np(base).copy()
        for m, o in enumerate(outs):
            acc = np.zeros(o.numel())
            for k, v in enumerate(vals):
                acc += W[m, k] * v
            if b is not None:
                acc += b
            if adds is not None and adds[m] is not None:
                acc += _np(adds[m])
            _np(o)[:] = acc

    @staticmethod
    def _node_values(ins, ncomp):
        vals = [_np(t).copy() for t in ins]
        return vals if ncomp == 1 else [vals[2 * j] + vals[2 * j + 1] for j in range(len(vals) // 2)]

    def colloc_sweep(self, ins, ncomp, outs, Wq=None, Wi=None, We=None, dt2=0.0, base=None, adds=None,
                     base_first=False):
        """Same order of operations as sdcb200_colloc_sweep (include/sdc_b200.h)."""
        self.launches += 1
        nj = len(ins) // ncomp
        raw = [_np(t).copy() for t in ins]
        F = self._node_values(ins, ncomp)
        b = None if base is None else _np(base).copy()
        shape = lambda W: np.asarray(W, dtype=float).reshape(len(outs), nj)  # noqa: E731
        for m, o in enumerate(outs):
            acc = b.copy() if base_first else np.zeros(o.numel())
            if Wq is not None:
                for j in range(nj):
                    acc += shape(Wq)[m, j] * F[j]
            if Wi is not None:
                for j in range(nj):
                    if ncomp == 1:
                        acc += shape(Wi)[m, j] * raw[j]
                    else:
                        acc += dt2 * (shape(Wi)[m, j] * raw[2 * j] + shape(We)[m, j] * raw[2 * j + 1])
            if b is not None and not base_first:
                acc += b
            if adds is not None and adds[m] is not None:
                acc += _np(adds[m])
            _np(o)[:] = acc

    def colloc_residual(self, Wq, ins, ncomp, u0, us, taus, res_outs, resnorm):
        self.launches += 1
        F = self._node_values(ins, ncomp)
        Wq = np.asarray(Wq, dtype=float).reshape(len(us), len(F))
        for m, u in enumerate(us):
            acc = np.zeros(u.numel())
            for j, v in enumerate(F):
                acc += Wq[m, j] * v
            acc += _np(u0) - _np(u)
            if taus is not None and taus[m] is not None:
                acc += _np(taus[m])
            if res_outs is not None and res_outs[m] is not None:
                _np(res_outs[m])[:] = acc
            resnorm[m] = float(np.max(np.abs(acc)))

    # ---- K2 ---------------------------------------------------------------------------------------------------------
    # ---- slabs: numpy views that include the two halo planes ----------------------------------------------------------
    @staticmethod
    def _slab_full(lay, t):
        """(nz+2, n, n) view of a slab field: plane 0 = lower halo, 1..nz owned, nz+1 = upper halo."""
        sz = lay.P * lay.P
        full = t.as_strided((sz * (lay.nz + 2),), (1,), t.storage_offset() - sz)
        return full.view(lay.nz + 2, lay.P, lay.P)[:, : lay.n, : lay.n]

    def _slab_lap_sum(self, xfull, periodic):
        """Neighbour sum on the owned planes of a slab: in-plane like _lap_sum, across planes from the halos."""
        x = xfull[1:-1]
        out = xfull[:-2] + xfull[2:]
        for ax in (1, 2):
            if periodic:
                out = out + np.roll(x, 1, axis=ax) + np.roll(x, -1, axis=ax)
            else:
                add = np.zeros_like(x)
                lo = [slice(None)] * 3
                hi = [slice(None)] * 3
                lo[ax], hi[ax] = slice(1, None), slice(None, -1)
                add[tuple(lo)] += x[tuple(hi)]
                add[tuple(hi)] += x[tuple(lo)]
                out = out + add
        return out

    def slab_cg_workspace(self, lay, comm, B):
        return None

    # ---- higher-order stencils: dense 1-D operators applied axis by axis -----------------------------------------------
    @staticmethod
    def _ho_matrix(op, n, periodic):
        h, order = op["order"] // 2, op["order"]
        A = np.zeros((n, n))
        for i in range(n):
            for k in range(-h, h + 1):
                j = i + k
                if periodic:
                    A[i, j % n] += op["centre"][abs(k)]
                elif 0 <= j < n:
                    A[i, j] += op["centre"][abs(k)]
        if not periodic:
            for i in range(h):
                A[i, :] = 0.0
                A[i, : order + 1] = op["lo"][i]
                A[n - 1 - i, :] = 0.0
                A[n - 1 - i, n - order - 1:] = op["hi"][i]
        return A

    def _ho_apply(self, op, lay, bc, x):
        A = self._ho_matrix(op, lay.n, bc == 1)
        out = np.zeros_like(x)
        for ax in range(x.ndim):
            out += np.moveaxis(np.tensordot(A, x, axes=([1], [ax])), 0, ax)
        return out

    def heat_eval_f_ho(self, lay, bc, op, us, fs, profile=None, gts=None, fexpls=None):
        self.launches += 1
        for i, (u, f) in enumerate(zip(us, fs)):
            self._grid(lay, f)[...] = self._ho_apply(op, lay, bc, self._grid(lay, u))
            if profile is not None:
                self._grid(lay, fexpls[i])[...] = self._grid(lay, profile) * gts[i]

    def cg_ho_workspace(self, lay, B):
        return None

    # ---- general finite-difference operators, GMRES -------------------------------------------------------------------
    @staticmethod
    def _fd_matrix(op, n, periodic):
        h = op["h"]
        A = np.zeros((n, n))
        for i in range(n):
            for k in range(-h, h + 1):
                j = i + k
                if periodic:
                    A[i, j % n] += op["coef"][k + h]
                elif 0 <= j < n:
                    A[i, j] += op["coef"][k + h]
        if not periodic:
            w = 2 * h + 1
            for i in range(h):
                A[i, :] = 0.0
                A[i, :w] = op["lo"][i]
                A[n - 1 - i, :] = 0.0
                A[n - 1 - i, n - w:] = op["hi"][i]
        return A

    def _fd_apply(self, op, lay, bc, x):
        A = self._fd_matrix(op, lay.n, bc == 1)
        out = np.zeros_like(x)
        for ax in range(x.ndim):
            out += np.moveaxis(np.tensordot(A, x, axes=([1], [ax])), 0, ax)
        return out

    def fd_eval_f(self, lay, bc, op, us, fs):
        self.launches += 1
        for u, f in zip(us, fs):
            self._grid(lay, f)[...] = self._fd_apply(op, lay, bc, self._grid(lay, u))

    def fd_gmres_workspace(self, lay, restart):
        return None

    @staticmethod
    def _lartg(f, g):
        """LAPACK dlartg as restated in csrc/gmres.cu (unscaled branch)."""
        if g == 0.0:
            return 1.0, 0.0, f
        if f == 0.0:
            return 0.0, np.copysign(1.0, g), abs(g)
        d = np.sqrt(f * f + g * g)
        r = np.copysign(d, f)
        return abs(f) / d, g / r, r

    def fd_gmres_solve(self, lay, bc, op, factor, rhs, x, rtol, maxiter, restart, work, iters_dev):
        """numpy mirror of csrc/gmres.cu::fd_gmres_kernel, statement by statement (same fused Gram-Schmidt passes, same
        scalar arithmetic), so that the CPU suite checks the device recurrence against the reference's fixtures."""
        self.launches += 1
        xg, b = self._grid(lay, x), self._grid(lay, rhs)
        eps = np.finfo(float).eps
        Mv = lambda v: v - factor * self._fd_apply(op, lay, bc, v)  # noqa: E731
        dot = lambda u, v: float(np.vdot(u, v))  # noqa: E731
        bnrm2 = np.sqrt(dot(b, b))
        if bnrm2 == 0.0:
            xg[...] = b
            return
        atol = rtol * bnrm2
        ptol_max_factor = 1.0
        ptol = bnrm2 * min(ptol_max_factor, atol / bnrm2)
        presid, inner_iter = 0.0, 0
        V = np.zeros((restart + 1,) + xg.shape)
        V[0] = b - Mv(xg)
        rnorm = np.sqrt(dot(V[0], V[0]))
        if rnorm < atol:
            return
        h = np.zeros((restart, restart + 1))
        giv = np.zeros((restart, 2))
        for iteration in range(maxiter):
            V[0] *= 1.0 / rnorm
            S = np.zeros(restart + 1)
            S[0] = rnorm
            breakdown = False
            for col in range(restart):
                w = V[col + 1]
                w[...] = Mv(V[col])
                h0 = np.sqrt(dot(w, w))
                hprev = 0.0
                for k in range(col + 1):
                    if k > 0:
                        w -= hprev * V[k - 1]
                    hprev = dot(V[k], w)
                    h[col, k] = hprev
                w -= hprev * V[col]
                h1 = np.sqrt(dot(w, w))
                if h1 <= eps * h0:
                    breakdown = True
                else:
                    w *= 1.0 / h1
                h[col, col + 1] = 0.0 if breakdown else h1
                for k in range(col):
                    c, s_ = giv[k]
                    n0, n1 = h[col, k], h[col, k + 1]
                    h[col, k], h[col, k + 1] = c * n0 + s_ * n1, -s_ * n0 + c * n1
                c, s_, mag = self._lartg(h[col, col], h[col, col + 1])
                giv[col] = c, s_
                h[col, col], h[col, col + 1] = mag, 0.0
                tmp = -s_ * S[col]
                S[col], S[col + 1] = c * S[col], tmp
                presid = abs(tmp)
                inner_iter += 1
                if inner_iter == maxiter:
                    break
                if presid <= ptol or breakdown:
                    break
            if h[col, col] == 0.0:
                S[col] = 0.0
            y = S[: col + 1].copy()
            for k in range(col, 0, -1):
                if y[k] != 0.0:
                    y[k] /= h[k, k]
                    y[:k] -= y[k] * h[k, :k]
            if y[0] != 0.0:
                y[0] /= h[0, 0]
            xg += np.tensordot(y, V[: col + 1], axes=1)
            V[0] = b - Mv(xg)
            rnorm = np.sqrt(dot(V[0], V[0]))
            if inner_iter == maxiter or rnorm <= atol or breakdown:
                break
            if presid <= ptol:
                ptol_max_factor = max(eps, 0.25 * ptol_max_factor)
            else:
                ptol_max_factor = min(1.0, 1.5 * ptol_max_factor)
            ptol = presid * min(ptol_max_factor, atol / rnorm)
        iters_dev[0] += inner_iter

    def heat_cg_solve_ho(self, lay, bc, op, factors, rhs, xs, rtol, maxiter, work, iters_dev):
        self.launches += 1
        for b, (r, x) in enumerate(zip(rhs, xs)):
            mv = lambda v, fac=factors[b]: v - fac * self._ho_apply(op, lay, bc, v)  # noqa: E731
            iters_dev[b] += self._cg(mv, self._grid(lay, r), self._grid(lay, x), rtol, maxiter)

    def heat_cg_solve_slab(self, lay, comm, bc, m_diag, m_off, rhs, xs, rtol, maxiter, work, iters_dev, precond=0):
        """Distributed CG with the same structure as the device solver: halo exchange of the search direction, global
        dot products through the communicator; the halo planes of xs are valid on entry."""
        from pysdc_b200.comm import SUM

        self.launches += 1
        nz = lay.nz
        for b, (r_t, x_t) in enumerate(zip(rhs, xs)):
            bvec = self._grid(lay, r_t).copy()
            x = self._grid(lay, x_t)
            dot = lambda u, v: comm.allreduce(float(np.vdot(u, v)), op=SUM)  # noqa: E731
            bb = dot(bvec, bvec)
            if bb == 0.0:
                x[...] = 0.0
                continue
            atol = rtol * np.sqrt(bb)
            xfull = _np(self._slab_full(lay, x_t))
            r = bvec - (m_diag[b] * xfull[1:-1] + m_off[b] * self._slab_lap_sum(xfull, bc == 1))
            pfull_t = torch.zeros((nz + 2, lay.n, lay.n), dtype=torch.float64)
            pfull = pfull_t.numpy()
            rfull_t = torch.zeros((nz + 2, lay.n, lay.n), dtype=torch.float64)
            rfull = rfull_t.numpy()
            cheb = self._cheb1(3, m_diag[b], m_off[b]) if precond else None
            rho_prev, its = None, 0
            for it in range(maxiter):
                if np.sqrt(dot(r, r)) < atol:
                    break
                if cheb is None:
                    z = r
                else:
                    rfull[1:-1] = r
                    comm.exchange_planes([(rfull_t[nz], rfull_t[0], rfull_t[1], rfull_t[nz + 1])], periodic=bc == 1)
                    z = cheb[0] * r + cheb[1] * (m_diag[b] * r + m_off[b] * self._slab_lap_sum(rfull, bc == 1))
                rho = dot(r, z)
                pfull[1:-1] = z if it == 0 else z + (rho / rho_prev) * pfull[1:-1]
                comm.exchange_planes([(pfull_t[nz], pfull_t[0], pfull_t[1], pfull_t[nz + 1])], periodic=bc == 1)
                q = m_diag[b] * pfull[1:-1] + m_off[b] * self._slab_lap_sum(pfull, bc == 1)
                alpha = rho / dot(pfull[1:-1], q)
                x += alpha * pfull[1:-1]
                r -= alpha * q
                rho_prev = rho
                its += 1
            iters_dev[b] += its

    def heat_eval_f(self, lay, bc, a_diag, a_off, us, fs, profile=None, gts=None, fexpls=None):
        self.launches += 1
        for i, (u, f) in enumerate(zip(us, fs)):
            if lay.is_slab:
                xfull = _np(self._slab_full(lay, u))
                self._grid(lay, f)[...] = a_diag * xfull[1:-1] + a_off * self._slab_lap_sum(xfull, bc == 1)
                if profile is not None:
                    self._grid(lay, fexpls[i])[...] = self._grid(lay, profile) * gts[i]
                continue
            x = self._grid(lay, u)
            self._grid(lay, f)[...] = a_diag * x + a_off * self._lap_sum(x, bc == 1)
            if profile is not None:
                self._grid(lay, fexpls[i])[...] = self._grid(lay, profile) * gts[i]

    def allencahn_eval_f(self, lay, a_diag, a_off, inv_eps2, nu_exp, us, fs, fexpls=None, split=None):
        self.launches += 1
        for i, (u, f) in enumerate(zip(us, fs)):
            x = self._grid(lay, u)
            lap = a_diag * x + a_off * self._lap_sum(x, True)
            react = inv_eps2 * x * (1.0 - x**nu_exp)
            if split == 2:
                self._grid(lay, f)[...] = lap - inv_eps2 * x ** (nu_exp + 1)
                self._grid(lay, fexpls[i])[...] = inv_eps2 * x
            elif fexpls is None:
                self._grid(lay, f)[...] = lap + react
            else:
                self._grid(lay, f)[...] = lap
                self._grid(lay, fexpls[i])[...] = react

    # ---- K5 ---------------------------------------------------------------------------------------------------------
    def upload_operator(self, W, col):
        return (torch.from_numpy(np.ascontiguousarray(W, dtype=np.float64)),
                torch.from_numpy(np.ascontiguousarray(col, dtype=np.int32)))

    def axis_apply(self, op, n_outer, n_out, n_inner, src, src_so, src_sa, dst, dst_so, dst_sa):
        self.launches += 1
        W, col = _np(op[0]), _np(op[1])
        s, d = _np(src), _np(dst)
        c = np.arange(n_inner)
        for o in range(n_outer):
            for i in range(n_out):
                acc = np.zeros(n_inner)
                for t in range(W.shape[1]):
                    if col[i, t] >= 0:
                        acc += W[i, t] * s[o * src_so + col[i, t] * src_sa + c]
                d[o * dst_so + i * dst_sa + c] = acc

    # ---- K3 / K4 ----------------------------------------------------------------------------------------------------
    def cg_workspace(self, lay, B):
        return torch.zeros(8, dtype=torch.float64)

    newton_workspace = lambda self, lay, B: torch.zeros(8, dtype=torch.float64)  # noqa: E731

    @staticmethod
    def _cheb1(ndim, m_diag, m_off):
        """Coefficients (a, b) of the degree-1 Chebyshev preconditioner z = a r + b M r (csrc/cg.cu::chebyshev1)."""
        w = 2.0 * ndim * abs(m_off)
        lmin, lmax = m_diag - w, m_diag + w
        theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
        if not delta > 0 or not lmin > 0:
            return 1.0 / theta, 0.0
        sigma = theta / delta
        rho0 = 1.0 / sigma
        rho1 = 1.0 / (2.0 * sigma - rho0)
        return (1.0 + rho1 * rho0) / theta + 2.0 * rho1 / delta, -2.0 * rho1 / (delta * theta)

    def _cg(self, matvec, b, x, rtol, maxiter, cheb=None):
        """scipy.sparse.linalg.cg's recurrence (atol = rtol*||b||, test before each iteration), optionally
        preconditioned with z = a r + b M r."""
        bnrm = np.linalg.norm(b)
        if bnrm == 0:
            x[...] = b
            return 0
        atol = rtol * bnrm
        r = b - matvec(x)
        p, rho_prev, its = None, None, 0
        for it in range(maxiter):
            if np.linalg.norm(r) < atol:
                return its
            z = r if cheb is None else cheb[0] * r + cheb[1] * matvec(r)
            rho = np.vdot(r, z)
            p = z.copy() if it == 0 else z + (rho / rho_prev) * p
            q = matvec(p)
            alpha = rho / np.vdot(p, q)
            x += alpha * p
            r -= alpha * q
            rho_prev = rho
            its += 1
        return its

    def heat_cg_solve(self, lay, bc, m_diag, m_off, rhs, xs, rtol, maxiter, work, iters_dev, precond=0):
        self.launches += 1
        for b, (r, x) in enumerate(zip(rhs, xs)):
            mv = lambda v, b=b: m_diag[b] * v + m_off[b] * self._lap_sum(v, bc == 1)  # noqa: E731
            cheb = self._cheb1(lay.ndim, m_diag[b], m_off[b]) if precond else None
            iters_dev[b] += self._cg(mv, self._grid(lay, r).copy(), self._grid(lay, x), rtol, maxiter, cheb)

    def heat_direct_solve_1d(self, lay, bc, m_diag, m_off, rhs, xs):
        self.launches += 1
        n = lay.n
        for b, (r, x) in enumerate(zip(rhs, xs)):
            Mx = m_diag[b] * np.eye(n) + m_off[b] * (np.eye(n, k=1) + np.eye(n, k=-1))
            if bc == 1:
                Mx[0, -1] += m_off[b]
                Mx[-1, 0] += m_off[b]
            self._grid(lay, x)[...] = np.linalg.solve(Mx, self._grid(lay, r))

    def reaction_workspace(self):
        return torch.zeros(1, dtype=torch.float64)

    def allencahn_reaction_newton(self, factors, inv_eps2, nu_exp, rhs, us, newton_tol, newton_maxiter, work, counters_dev):
        """sdcb200_allencahn_reaction_newton: global Newton loop, diagonal Jacobian solved exactly."""
        self.launches += 1
        for factor, r, u in zip(factors, rhs, us):
            x, b = _np(u), _np(r)
            n = 0
            while n < newton_maxiter:
                g = x - factor * (inv_eps2 * x * (1.0 - x**nu_exp)) - b
                if np.max(np.abs(g)) < newton_tol:
                    break
                x -= g / (1.0 - factor * (inv_eps2 * (1.0 - (nu_exp + 1) * x**nu_exp)))
                n += 1
            counters_dev[0] += n

    def allencahn_newton_solve(self, lay, factors, a_diag, a_off, inv_eps2, nu_exp, rhs, us, newton_tol, newton_maxiter,
                               lin_tol, lin_maxiter, inexact_ratio, work, counters_dev, variant=0):
        self.launches += 1
        for factor, r, u in zip(factors, rhs, us):
            x = self._grid(lay, u)
            b = self._grid(lay, r)
            n, tol = 0, lin_tol
            while n < newton_maxiter:
                lap = a_diag * x + a_off * self._lap_sum(x, True)
                if variant == 0:
                    g = x - factor * (lap + inv_eps2 * x * (1.0 - x**nu_exp)) - b
                else:
                    g = x - factor * (lap - inv_eps2 * x ** (nu_exp + 1)) - b
                res = np.max(np.abs(g))
                if inexact_ratio:
                    tol = res * inexact_ratio
                if res < newton_tol:
                    break
                if variant == 0:
                    d = 1.0 - factor * (a_diag + inv_eps2 * (1.0 - (nu_exp + 1) * x**nu_exp))
                else:
                    d = 1.0 - factor * (a_diag - inv_eps2 * ((nu_exp + 1) * x**nu_exp))
                mv = lambda v: d * v - factor * a_off * self._lap_sum(v, True)  # noqa: E731
                z = np.zeros_like(x)
                counters_dev[1] += self._cg(mv, g, z, tol, lin_maxiter)
                x -= z
                n += 1
            counters_dev[0] += n
