"""Problem classes of the sweep path with the reference's constructor keywords, attributes and error behaviour:

* ``heatNd_unforced`` / ``heatNd_forced`` — ``pySDC/implementations/problem_classes/HeatEquation_ND_FD.py:8-230`` on top of
  ``GenericNDimFinDiff`` (``generic_ND_FD.py:13-264``),
* ``allencahn_fullyimplicit`` — ``problem_classes/AllenCahn_2D_FD.py:14-257``.

The spatial operator is never assembled (the reference builds a scipy sparse matrix, 934 M non-zeros at 511^3): ``eval_f``
is the matrix-free stencil kernel and ``solve_system`` the persistent CG / Newton kernels of ``libsdcb200``.  Besides the
reference API the classes offer batched, in-place entry points (``eval_f_batch``, ``solve_system_batch``) that the
sweepers use to update all M nodes with one launch when QDelta is diagonal.

Classes are written as mix-ins; ``_bind(Problem)`` attaches them to a base class — the stand-alone one from
``core.py`` here, pySDC's own in ``pysdc_plugin.py``.
"""
import numpy as np
import torch

from .backend import get_backend
from .datatypes import comp2_mesh, imex_mesh, mesh
from .layout import get_layout
from .errors import ProblemError
from .fields_io import OutputMixin

BC_CODES = {"dirichlet-zero": 0, "periodic": 1}
GMRES_RESTART = 20  # scipy's default restart length, the only one the reference uses (generic_ND_FD.py:241-250)


def grid_1d(size, bc, left=0.0, right=1.0):
    """Mesh width and 1-D grid (helpers/problem_helper.py:245-269)."""
    L = right - left
    if bc == "periodic":
        dx = L / size
        x = np.array([left + dx * i for i in range(size)])
    elif "dirichlet" in bc or "neumann" in bc:
        dx = L / (size + 1)
        x = np.array([left + dx * (i + 1) for i in range(size)])
    else:
        raise NotImplementedError(f'Boundary conditions "{bc}" not implemented.')
    return dx, x


def fd_weights(derivative, offsets):
    """Finite-difference weights on the given integer offsets from the Taylor system, solved the way the reference
    solves it (helpers/problem_helper.py:42-80) so that the coefficients are the reference's floats."""
    from scipy.special import factorial

    steps = np.asarray(offsets)
    n = len(steps)
    A = np.zeros((n, n))
    idx = np.arange(n)
    inv_facs = 1.0 / factorial(idx)
    for i in range(n):
        A[i, :] = steps ** idx[i] * inv_facs[i]
    sol = np.zeros(n)
    sol[derivative] = 1.0
    coeff = np.linalg.solve(A, sol)
    return coeff[np.argsort(steps)], np.sort(steps)


def high_order_tables(order, bc, scale):
    """Coefficient tables of the centred second-derivative stencil of the given order, times ``scale`` (= nu / dx^2):
    ``centre[k]``, k = 0 .. order/2, and - dirichlet-zero - the one-sided closure rows of the order/2 points next to
    each boundary (helpers/problem_helper.py:157-201 with the default ``reduce=False``; the coefficient of the boundary
    value itself drops out because that value is zero): ``lo[i]`` acts on columns 0 .. order, ``hi[i]`` (row n-1-i) on
    the last order+1 columns."""
    h = order // 2
    w, _ = fd_weights(2, np.arange(order + 1) - h)
    tables = dict(order=order, centre=[float(w[h + k]) * scale for k in range(h + 1)], lo=None, hi=None)
    if bc != "periodic":
        lo, hi = np.zeros((h, order + 1)), np.zeros((h, order + 1))
        for i in range(h):
            cl, _ = fd_weights(2, np.arange(-(i + 1), order + 2 - (i + 1)))
            cr, _ = fd_weights(2, np.arange(-(order + 2) + (i + 2), (i + 2)))
            lo[i] = cl[1:] * scale
            hi[i] = cr[:-1] * scale
        tables.update(lo=lo, hi=hi)
    return tables


def fd_steps(derivative, order, stencil_type):
    """Offsets of the stencil (helpers/problem_helper.py:4-39)."""
    if stencil_type == "center":
        n = order + derivative - (derivative + 1) % 2 // 1
        return np.arange(n) - n // 2
    if stencil_type == "forward":
        return np.arange(order + derivative)
    if stencil_type == "backward":
        return -np.arange(order + derivative)
    if stencil_type == "upwind":
        n = order + derivative
        return -np.arange(n) if n <= 3 else np.append(-np.arange(n - 1)[::-1], [1])
    raise ValueError(f'Stencil must be of type "center", "forward", "backward" or "upwind", not {stencil_type}.')


def fd_operator_tables(derivative, order, stencil_type, bc, dx, coeff):
    """The 1-D operator ``coeff * d^derivative/dx^derivative`` of the reference (helpers/problem_helper.py:83-242,
    generic_ND_FD.py:140-149) as what the device needs: half width ``h``, ``coef[k + h]`` = coefficient of offset k, and
    on dirichlet-zero grids the first / last h rows of the matrix (``lo[i]``: row i on columns 0 .. 2h; ``hi[i]``: row
    n-1-i on the last 2h+1 columns), which hold the reference's one-sided closure stencils (default ``reduce=False``;
    the coefficient of the zero boundary value drops out) and, where a one-sided stencil needs no closure on that side,
    the truncated interior stencil."""
    w, steps = fd_weights(derivative, fd_steps(derivative, order, stencil_type))
    h = int(max(-steps.min(), steps.max()))
    scale = lambda v: (np.asarray(v, dtype=float) / dx**derivative) * coeff  # noqa: E731  (the reference's order)
    coef = np.zeros(2 * h + 1)
    coef[steps + h] = w
    tables = dict(h=h, coef=scale(coef), lo=None, hi=None)
    if bc != "periodic":
        m = 4 * h + 2  # any size with the two boundary blocks apart gives the same rows
        A = np.zeros((m, m))
        for c, k in zip(w, steps):
            A += c * np.eye(m, k=int(k))
        for side, width in ((0, int(-steps.min())), (1, int(steps.max()))):
            for i in range(width):
                if side == 0:
                    bw, _ = fd_weights(derivative, np.arange(-(i + 1), order + derivative - (i + 1)))
                    A[i, :] = 0.0
                    A[i, : len(bw) - 1] = bw[1:]
                else:
                    bw, _ = fd_weights(derivative, np.arange(-(order + derivative) + (i + 2), (i + 2)))
                    A[-i - 1, :] = 0.0
                    A[-i - 1, -len(bw) + 1:] = bw[:-1]
        if order + derivative - 1 > 2 * h + 1:
            raise ProblemError("closure rows wider than the device stencil tables")
        tables.update(lo=scale(A[:h, : 2 * h + 1]), hi=scale(A[::-1][:h, -(2 * h + 1):]))
    return tables


def sparse_operator(tables, n, ndim, periodic):
    """The spatial operator as a scipy sparse matrix on the host - the attribute ``A`` of the reference's problem classes
    (generic_ND_FD.py:140-149; Kronecker sum of helpers/problem_helper.py:226-239), assembled from the same 1-D tables
    the kernels use.  Set-up / analysis aid only (tutorial/step_1/C builds the collocation matrix from it); the sweep
    never touches it."""
    import scipy.sparse as sp

    h, coef = tables["h"], np.asarray(tables["coef"], dtype=float)
    A1 = sp.lil_matrix((n, n))
    for i in range(n):
        if not periodic and i < h:
            row = np.asarray(tables["lo"])[i]
            for j in range(min(2 * h + 1, n)):
                A1[i, j] = row[j]
        elif not periodic and i >= n - h:
            row = np.asarray(tables["hi"])[n - 1 - i]
            for j in range(min(2 * h + 1, n)):
                A1[i, n - (2 * h + 1) + j] = row[j]
        else:
            for k in range(-h, h + 1):
                if coef[k + h] != 0.0:
                    A1[i, (i + k) % n] += coef[k + h]
    A1 = A1.tocsc()
    if ndim == 1:
        return A1
    if ndim == 2:
        return (sp.kron(A1, sp.eye(n)) + sp.kron(sp.eye(n), A1)).tocsc()
    return (sp.kron(A1, sp.eye(n**2)) + sp.kron(sp.eye(n**2), A1) + sp.kron(sp.kron(sp.eye(n), A1), sp.eye(n))).tocsc()


class _HostOperator:
    """``A`` and ``Id`` of the reference's finite-difference problem classes, built on first use."""

    def _operator_tables(self):
        raise NotImplementedError

    @property
    def A(self):
        if "_A_host" not in self.__dict__:
            self.__dict__["_A_host"] = sparse_operator(self._operator_tables(), self.nvars[0], len(self.nvars),
                                                       self.bc == "periodic")
        return self.__dict__["_A_host"]

    @property
    def Id(self):
        import scipy.sparse as sp

        return sp.eye(int(np.prod(self.nvars)), format="csc")


def check_fd_params(nvars, freq, bc):
    """Parameter checks of GenericNDimFinDiff (generic_ND_FD.py:99-133); returns (nvars, freq, ndim, bc)."""
    if type(nvars) not in [int, tuple]:
        raise ProblemError("nvars should be either tuple or int")
    if type(freq) not in [int, tuple]:
        raise ProblemError("freq should be either tuple or int")
    if type(nvars) is int:
        nvars = (nvars,)
    ndim = len(nvars)
    if ndim > 3:
        raise ProblemError(f"can work with up to three dimensions, got {ndim}")
    if type(freq) is int:
        freq = (freq,) * ndim
    if len(freq) != ndim:
        raise ProblemError(f"len(freq)={len(freq)}, different to ndim={ndim}")
    for f in freq:
        if ndim == 1 and f == -1:
            bc = "periodic"
            break
        if f % 2 != 0 and bc == "periodic":
            raise ProblemError("need even number of frequencies due to periodic BCs")
    for nvar in nvars:
        if nvar % 2 != 0 and bc == "periodic":
            raise ProblemError("the setup requires nvars = 2^p per dimension")
        if (nvar + 1) % 2 != 0 and bc == "dirichlet-zero":
            raise ProblemError("setup requires nvars = 2^p - 1")
    if ndim > 1 and nvars[1:] != nvars[:-1]:
        raise ProblemError("need a square domain, got %s" % (nvars,))
    return nvars, freq, ndim, bc


class DeviceWorkCounter:
    """``WorkCounter`` (core/problem.py:16-40) whose count lives in a device int: the solver kernels add their
    iteration counts without a host round trip; reading ``niter`` synchronises."""

    def __init__(self, slot):
        self._slot = slot  # 1-element int32 device tensor
        self._host = 0

    def __call__(self, *args, **kwargs):
        self._host += 1

    def decrement(self):
        self._host -= 1

    @property
    def niter(self):
        return int(self._slot.item()) + self._host

    def __str__(self):
        return f"{self.niter}"


# ---------------------------------------------------------------------------------------------------------------------
# heat equation
# ---------------------------------------------------------------------------------------------------------------------
class HeatMixin(_HostOperator, OutputMixin):
    dtype_u = mesh
    dtype_f = mesh
    forced = False

    def __init__(self, nvars=512, nu=0.1, freq=2, stencil_type="center", order=2, lintol=1e-12, liniter=10000,
                 solver_type="direct", bc="periodic", sigma=6e-2, comm=None, preconditioner=None):
        # `comm` is the one keyword the reference does not have (its FD problems are not space-parallel): a
        # parallel.SlabComm decomposes the 3-D grid into slabs along axis 0, one per GPU; nvars stays the GLOBAL shape.
        # `preconditioner='chebyshev'` (2-D / 3-D dirichlet-zero, CG) preconditions the node solves with a degree-1
        # Chebyshev polynomial of the operator: same stopping test and tolerance as the reference's plain CG, about
        # half the iterations (work_counters['CG'] then counts preconditioned iterations).
        # parameter checks of generic_ND_FD.py:99-133
        nvars, freq, ndim, bc = check_fd_params(nvars, freq, bc)
        # what the device path implements
        if bc not in BC_CODES:
            raise ProblemError(f"boundary condition {bc!r} is not implemented on the device (have {list(BC_CODES)})")
        if order not in (2, 4, 6, 8) or stencil_type != "center":
            raise ProblemError("the device stencils are the centred Laplacians of order 2, 4, 6 and 8; "
                               f"got order={order}, stencil_type={stencil_type!r}")
        if order != 2 and (preconditioner is not None or comm is not None):
            raise ProblemError("order > 2 is implemented without preconditioner and on one GPU")
        if order != 2 and any(nv <= order for nv in nvars):
            raise ProblemError(f"grid too small for the order-{order} stencil")
        if solver_type not in ("CG", "GMRES", "direct"):
            raise ProblemError(f"solver_type {solver_type!r} is not implemented on the device (have 'CG', 'GMRES', "
                               "'direct')")
        if solver_type == "GMRES" and (preconditioner is not None or comm is not None):
            raise ProblemError("solver_type='GMRES' runs without preconditioner and on one GPU")
        # solver_type='direct' (the reference's default): a device factorisation exists for 1-D order-2 grids of 3 .. 8192
        # points (tridiagonal Thomas solve in shared memory); every other grid runs the CG down to its attainable accuracy instead (solve_system_batch)
        self._direct_ok = ndim == 1 and order == 2 and 3 <= nvars[0] <= 8192  # (sdcb200_heat_direct_solve_1d's range)

        if preconditioner not in (None, "chebyshev"):
            raise ProblemError(f"unknown preconditioner {preconditioner!r} (have None, 'chebyshev')")
        if preconditioner is not None and (solver_type != "CG" or bc != "dirichlet-zero" or ndim < 2):
            raise ProblemError("the polynomial preconditioner needs solver_type='CG' on a 2-D / 3-D dirichlet-zero grid")
        slab = comm is not None and hasattr(comm, "slab_layout")
        if slab and (ndim != 3 or solver_type != "CG" or bc != "dirichlet-zero"):
            raise ProblemError("slab-decomposed runs are implemented for 3-D dirichlet-zero grids with solver_type='CG'")

        super().__init__(init=(nvars[0] if ndim == 1 else nvars, comm, np.dtype("float64")))

        dx, xvalues = grid_1d(nvars[0], bc)
        self.xvalues = xvalues
        self._makeAttributeAndRegister("nvars", "stencil_type", "order", "bc", localVars=locals(), readOnly=True)
        self._makeAttributeAndRegister("freq", "lintol", "liniter", "solver_type", localVars=locals())
        self._makeAttributeAndRegister("nu", localVars=locals(), readOnly=True)
        self._makeAttributeAndRegister("sigma", localVars=locals())

        # entries of A = nu * (kron-sum of [1, -2, 1]) / dx^2, formed in the order the reference forms them
        # (problem_helper.py:239 divides by dx**2, generic_ND_FD.py:149 multiplies by the coefficient)
        self.a_off = (1.0 / dx**2) * nu
        self.a_diag = ((-2.0 * ndim) / dx**2) * nu
        self._bc = BC_CODES[bc]
        self._be = get_backend()
        # order > 2: wide stencils with the reference's boundary closures (highorder.cu); a_diag stays the diagonal of A
        self._ho = None if order == 2 else high_order_tables(order, bc, (1.0 / dx**2) * nu)
        if self._ho is not None:
            self.a_diag = ndim * self._ho["centre"][0]
        # GMRES runs on the general operator tables (gmres.cu), whatever the order
        self._fd = fd_operator_tables(2, order, stencil_type, bc, dx, nu) if solver_type == "GMRES" else None
        self._precond = 1 if preconditioner == "chebyshev" else 0
        self._comm = comm if slab else None
        self._lay = comm.slab_layout(nvars) if slab else get_layout(nvars)
        self._counters = self._be.zeros(2 + 8, dtype=torch.int32)  # [total CG its, unused, per-system its of a solve]
        self._work = {}
        self._profile = None
        if solver_type != "direct":
            self.work_counters[solver_type] = DeviceWorkCounter(self._counters[0:1])

    # -- reference attributes -----------------------------------------------------------------------------------------
    def _operator_tables(self):
        return fd_operator_tables(2, self.order, self.stencil_type, self.bc, self.dx, self.nu)

    @property
    def ndim(self):
        return len(self.nvars)

    @property
    def dx(self):
        return self.xvalues[1] - self.xvalues[0]

    @property
    def grids(self):
        x = self.xvalues
        if self.ndim == 1:
            return x
        if self.ndim == 2:
            return x[None, :], x[:, None]
        return x[None, :, None], x[:, None, None], x[None, None, :]

    @classmethod
    def get_default_sweeper_class(cls):
        from .sweepers import generic_implicit

        return generic_implicit

    # -- right-hand side ----------------------------------------------------------------------------------------------
    def _spatial_profile(self):
        """prod_d sin(pi k_d x_d) evaluated on the host with the reference's numpy expression (HeatEquation_ND_FD.py:
        184-203) and uploaded once: the forcing is this profile times a scalar g(t), so no device sin is needed."""
        if self._profile is None:
            g = self.grids if self.ndim > 1 else (self.grids,)
            prof = np.sin(np.pi * self.freq[0] * g[0])
            for k, x in zip(self.freq[1:], g[1:]):
                prof = prof * np.sin(np.pi * k * x)
            m = mesh(self.init)
            m[:] = np.broadcast_to(prof, self.nvars)
            self._profile = m
        return self._profile

    def _forcing_factor(self, t):
        return self.nu * np.pi**2 * sum([k**2 for k in self.freq]) * np.cos(t) - np.sin(t)

    def eval_f_batch(self, us, ts, fs):
        """fs[i] = f(us[i], ts[i]) in place, one launch for all fields."""
        if self._ho is not None:
            args = (self._spatial_profile().flat, [self._forcing_factor(t) for t in ts],
                    [f.expl.flat for f in fs]) if self.forced else ()
            self._be.heat_eval_f_ho(self._lay, self._bc, self._ho, [u.flat for u in us],
                                    [(f.impl if self.forced else f).flat for f in fs], *args)
            return
        if self._comm is not None:
            self._comm.exchange_halos(us)
        if self.forced:
            self._be.heat_eval_f(self._lay, self._bc, self.a_diag, self.a_off, [u.flat for u in us],
                                 [f.impl.flat for f in fs], self._spatial_profile().flat,
                                 [self._forcing_factor(t) for t in ts], [f.expl.flat for f in fs])
        else:
            self._be.heat_eval_f(self._lay, self._bc, self.a_diag, self.a_off, [u.flat for u in us],
                                 [f.flat for f in fs])

    def eval_f(self, u, t):
        """generic_ND_FD.py:188-206 / HeatEquation_ND_FD.py:162-204."""
        f = self.f_init
        self.eval_f_batch([u], [t], [f])
        return f

    # -- implicit solves ----------------------------------------------------------------------------------------------
    def _cg_work(self, B):
        if B not in self._work:
            if self.solver_type == "GMRES":
                self._work[B] = self._be.fd_gmres_workspace(self._lay, GMRES_RESTART)
            elif self._ho is not None:
                self._work[B] = self._be.cg_ho_workspace(self._lay, B)
            elif self._comm is not None:
                import weakref

                self._work[B] = self._be.slab_cg_workspace(self._lay, self._comm, B)
                if hasattr(self._work[B], "close"):
                    weakref.finalize(self, self._work[B].close)  # unmap the peers' segments, free our own
            else:
                self._work[B] = self._be.cg_workspace(self._lay, B)
        return self._work[B]

    def solve_system_batch(self, rhs, factors, xs, ts=None):
        """Solve (I - factors[i] A) xs[i] = rhs[i] in place (xs[i] holds the initial guess), all systems in one
        persistent launch."""
        m_diag = [1.0 - f * self.a_diag for f in factors]
        m_off = [-(f * self.a_off) for f in factors]
        if self.solver_type == "direct" and self._direct_ok:
            self._be.heat_direct_solve_1d(self._lay, self._bc, m_diag, m_off, [r.flat for r in rhs],
                                          [x.flat for x in xs])
            return
        if self.solver_type == "direct":
            # no device factorisation beyond the 1-D order-2 tridiagonal one: the system is solved by the persistent CG
            # iterated down to the accuracy the floating-point recurrence can attain (rtol 1e-15 on the recurrence
            # residual, the same forward error ~ cond * eps a sparse direct solve leaves); like the reference's 'direct'
            # it counts no work (generic_ND_FD.py:236-239)
            scratch = self._counters[2: 2 + len(xs)]
            n_total = int(np.prod(self.nvars))
            if self._ho is not None:
                self._be.heat_cg_solve_ho(self._lay, self._bc, self._ho, list(factors), [r.flat for r in rhs],
                                          [x.flat for x in xs], 1e-15, 10 * n_total, self._cg_work(len(xs)), scratch)
            else:
                self._be.heat_cg_solve(self._lay, self._bc, m_diag, m_off, [r.flat for r in rhs], [x.flat for x in xs],
                                       1e-15, 10 * n_total, self._cg_work(len(xs)), scratch)
            return
        counters = self._counters[2: 2 + len(xs)]
        counters.zero_()
        log = getattr(self, "solve_log", None)  # bench.py: per-launch device timing + iteration counts
        if log is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        if self.solver_type == "GMRES":  # generic_ND_FD.py:241-250: one restarted-GMRES launch per system
            work = self._cg_work(1)
            for i, (f, r, x) in enumerate(zip(factors, rhs, xs)):
                self._be.fd_gmres_solve(self._lay, self._bc, self._fd, f, r.flat, x.flat, self.lintol, self.liniter,
                                        GMRES_RESTART, work, counters[i: i + 1])
        elif self._ho is not None:
            self._be.heat_cg_solve_ho(self._lay, self._bc, self._ho, list(factors), [r.flat for r in rhs],
                                      [x.flat for x in xs], self.lintol, self.liniter, self._cg_work(len(xs)), counters)
        elif self._comm is not None:
            # the initial guesses need their neighbours' boundary planes; the solver exchanges everything else itself
            work = self._cg_work(len(xs))
            self._comm.exchange_halos(xs)
            self._be.heat_cg_solve_slab(self._lay, self._comm, self._bc, m_diag, m_off, [r.flat for r in rhs],
                                        [x.flat for x in xs], self.lintol, self.liniter, work, counters,
                                        precond=self._precond)
        else:
            self._be.heat_cg_solve(self._lay, self._bc, m_diag, m_off, [r.flat for r in rhs], [x.flat for x in xs],
                                   self.lintol, self.liniter, self._cg_work(len(xs)), counters,
                                   precond=self._precond)
        if log is not None:
            ev1.record()
            log.append((ev0, ev1, counters.clone()))
        # the reference counts one callback per CG iteration of every solve
        self._counters[0:1] += counters.sum(dtype=torch.int32)

    def solve_system(self, rhs, factor, u0, t):
        """generic_ND_FD.py:208-264: returns a new field, inputs untouched."""
        sol = self.dtype_u(u0)
        self.solve_system_batch([rhs], [factor], [sol], [t])
        return sol

    # -- exact solutions ----------------------------------------------------------------------------------------------
    def u_exact(self, t, **kwargs):
        ndim, freq, nu, sigma, dx = self.ndim, self.freq, self.nu, self.sigma, self.dx
        sol = self.u_init
        if self.forced:  # HeatEquation_ND_FD.py:206-230
            g = self.grids if ndim > 1 else (self.grids,)
            val = np.sin(np.pi * freq[0] * g[0])
            for k, x in zip(freq[1:], g[1:]):
                val = val * np.sin(np.pi * k * x)
            sol[:] = np.broadcast_to(val * np.cos(t), self.nvars)
            return sol
        # HeatEquation_ND_FD.py:84-132
        if ndim == 1:
            x = self.grids
            rho = (2.0 - 2.0 * np.cos(np.pi * freq[0] * dx)) / dx**2
            if freq[0] > 0:
                sol[:] = np.sin(np.pi * freq[0] * x) * np.exp(-t * nu * rho)
            elif freq[0] == -1:
                sol[:] = np.exp(-0.5 * ((x - 0.5) / sigma) ** 2) * np.exp(-t * nu * rho)
        elif ndim == 2:
            rho = (2.0 - 2.0 * np.cos(np.pi * freq[0] * dx)) / dx**2 + (2.0 - 2.0 * np.cos(np.pi * freq[1] * dx)) / dx**2
            x, y = self.grids
            sol[:] = np.sin(np.pi * freq[0] * x) * np.sin(np.pi * freq[1] * y) * np.exp(-t * nu * rho)
        else:
            # the middle term lacks /dx**2 in the reference (:119-123); kept so that u_exact is identical
            rho = ((2.0 - 2.0 * np.cos(np.pi * freq[0] * dx)) / dx**2 + (2.0 - 2.0 * np.cos(np.pi * freq[1] * dx))
                   + (2.0 - 2.0 * np.cos(np.pi * freq[2] * dx)) / dx**2)
            x, y, z = self.grids
            sol[:] = (np.sin(np.pi * freq[0] * x) * np.sin(np.pi * freq[1] * y) * np.sin(np.pi * freq[2] * z)
                      * np.exp(-t * nu * rho))
        return sol


class HeatForcedMixin(HeatMixin):
    dtype_f = imex_mesh
    forced = True

    @classmethod
    def get_default_sweeper_class(cls):
        from .sweepers import imex_1st_order

        return imex_1st_order


# ---------------------------------------------------------------------------------------------------------------------
# advection equation
# ---------------------------------------------------------------------------------------------------------------------
class AdvectionMixin(_HostOperator, OutputMixin):
    """``advectionNd`` (problem_classes/AdvectionEquation_ND_FD.py:7-164 on top of ``GenericNDimFinDiff``): u_t = -c
    grad u in 1-3 dimensions, any stencil of helpers/problem_helper.py:4-39 within four grid points (centred order 2-8,
    upwind order 1-5, forward / backward order 1-4), ``eval_f`` by the general finite-difference kernel and the node
    solves by the device GMRES (``solver_type='GMRES'``; the non-symmetric systems have no device 'direct' or 'CG'
    path)."""

    dtype_u = mesh
    dtype_f = mesh
    forced = False

    def __init__(self, nvars=512, c=1.0, freq=2, stencil_type="center", order=2, lintol=1e-12, liniter=10000,
                 solver_type="direct", bc="periodic", sigma=6e-2):
        nvars, freq, ndim, bc = check_fd_params(nvars, freq, bc)
        if bc not in BC_CODES:
            raise ProblemError(f"boundary condition {bc!r} is not implemented on the device (have {list(BC_CODES)})")
        if solver_type not in ("GMRES", "direct", "CG"):
            raise ProblemError(f"solver_type {solver_type!r} is not implemented on the device for the advection "
                               "equation: its systems are non-symmetric, use solver_type='GMRES'")
        try:
            steps = fd_steps(1, order, stencil_type)
        except ValueError as e:
            raise ProblemError(str(e)) from None
        if max(-steps.min(), steps.max()) > 4 or order < 1:
            raise ProblemError(f"the device stencils reach at most four grid points; got order={order}, "
                               f"stencil_type={stencil_type!r}")
        if any(nv <= 2 * max(-steps.min(), steps.max()) for nv in nvars):
            raise ProblemError(f"grid too small for the order-{order} stencil")
        super().__init__(init=(nvars[0] if ndim == 1 else nvars, None, np.dtype("float64")))
        dx, xvalues = grid_1d(nvars[0], bc)
        self.xvalues = xvalues
        self._makeAttributeAndRegister("nvars", "stencil_type", "order", "bc", localVars=locals(), readOnly=True)
        self._makeAttributeAndRegister("freq", "lintol", "liniter", "solver_type", localVars=locals())
        self._makeAttributeAndRegister("c", localVars=locals(), readOnly=True)
        self._makeAttributeAndRegister("sigma", localVars=locals())
        self._bc = BC_CODES[bc]
        self._be = get_backend()
        self._lay = get_layout(nvars)
        self._fd = fd_operator_tables(1, order, stencil_type, bc, dx, -c)  # AdvectionEquation_ND_FD.py:89: coeff = -c
        self._counters = self._be.zeros(2, dtype=torch.int32)  # [total GMRES its, its of the current solve]
        self._work = None
        if solver_type != "direct":
            self.work_counters[solver_type] = DeviceWorkCounter(self._counters[0:1])

    def _operator_tables(self):
        return self._fd

    ndim = HeatMixin.ndim
    dx = HeatMixin.dx
    grids = HeatMixin.grids
    get_default_sweeper_class = HeatMixin.get_default_sweeper_class

    def eval_f_batch(self, us, ts, fs):
        self._be.fd_eval_f(self._lay, self._bc, self._fd, [u.flat for u in us], [f.flat for f in fs])

    def eval_f(self, u, t):
        """generic_ND_FD.py:188-206."""
        f = self.f_init
        self.eval_f_batch([u], [t], [f])
        return f

    def solve_system_batch(self, rhs, factors, xs, ts=None):
        """(I - factors[i] A) xs[i] = rhs[i] in place by restarted GMRES (generic_ND_FD.py:241-250), one persistent
        launch per system."""
        if self.solver_type == "CG":  # the problem can be set up and evaluated with it; the systems are non-symmetric
            raise ProblemError("solver_type 'CG' is not implemented on the device for the advection equation: its systems "
                               "are non-symmetric, use solver_type='GMRES'")
        if self._work is None:
            self._work = self._be.fd_gmres_workspace(self._lay, GMRES_RESTART)
        counter = self._counters[1:2]
        # solver_type='direct' (the reference's default): no device factorisation of these banded non-symmetric systems;
        # the restarted GMRES is iterated down to a relative residual of 1e-14 instead and, like the reference's
        # 'direct', counts no work
        direct = self.solver_type == "direct"
        rtol, cap = (1e-14, 50 * GMRES_RESTART) if direct else (self.lintol, self.liniter)
        for f, r, x in zip(factors, rhs, xs):
            self._be.fd_gmres_solve(self._lay, self._bc, self._fd, f, r.flat, x.flat, rtol, cap, GMRES_RESTART,
                                    self._work, counter)
        if not direct:
            self._counters[0:1] += counter
        counter.zero_()

    def solve_system(self, rhs, factor, u0, t):
        sol = self.dtype_u(u0)
        self.solve_system_batch([rhs], [factor], [sol], [t])
        return sol

    def u_exact(self, t, **kwargs):
        """AdvectionEquation_ND_FD.py:95-141."""
        ndim, freq, c, sigma, sol = self.ndim, self.freq, self.c, self.sigma, self.u_init
        if ndim == 1:
            x = self.grids
            if freq[0] >= 0:
                sol[:] = np.sin(np.pi * freq[0] * (x - c * t))
            elif freq[0] == -1:
                sol[:] = np.exp(-0.5 * (((x - (c * t)) % 1.0 - 0.5) / sigma) ** 2)
        elif ndim == 2:
            x, y = self.grids
            sol[:] = np.sin(np.pi * freq[0] * (x - c * t)) * np.sin(np.pi * freq[1] * (y - c * t))
        else:
            x, y, z = self.grids
            sol[:] = (np.sin(np.pi * freq[0] * (x - c * t)) * np.sin(np.pi * freq[1] * (y - c * t))
                      * np.sin(np.pi * freq[2] * (z - c * t)))
        return sol


# ---------------------------------------------------------------------------------------------------------------------
# Allen-Cahn, fully implicit
# ---------------------------------------------------------------------------------------------------------------------
class AllenCahnMixin(OutputMixin):
    dtype_u = mesh
    dtype_f = mesh
    forced = False

    def __init__(self, nvars=(128, 128), nu=2, eps=0.04, newton_maxiter=200, newton_tol=1e-12, lin_tol=1e-8,
                 lin_maxiter=100, inexact_linear_ratio=None, radius=0.25, order=2):
        if len(nvars) != 2:
            raise ProblemError("this is a 2d example, got %s" % (nvars,))
        if nvars[0] != nvars[1]:
            raise ProblemError("need a square domain, got %s" % (nvars,))
        if nvars[0] % 2 != 0:
            raise ProblemError("the setup requires nvars = 2^p per dimension")
        if order != 2:
            raise ProblemError(f"the device stencil is the order-2 centred Laplacian; got order={order}")
        if int(nu) != nu or nu < 1:
            raise ProblemError(f"the device path implements integer exponents nu >= 1, got {nu}")
        nvars = tuple(nvars)
        super().__init__((nvars, None, np.dtype("float64")))
        self._makeAttributeAndRegister("nvars", "nu", "eps", "radius", "order", localVars=locals(), readOnly=True)
        self._makeAttributeAndRegister("newton_maxiter", "newton_tol", "lin_tol", "lin_maxiter", "inexact_linear_ratio",
                                       localVars=locals(), readOnly=False)
        self.dx = 1.0 / self.nvars[0]
        self.xvalues = np.array([i * self.dx - 0.5 for i in range(self.nvars[0])])
        self.a_off = 1.0 / self.dx**2
        self.a_diag = (-2.0 * 2) / self.dx**2
        self._be = get_backend()
        self._lay = get_layout(nvars)
        # [work newton, work linear, rhs(unused), -, per-launch newton, per-launch linear, newton_itercount, lin_itercount]
        self._counters = self._be.zeros(8, dtype=torch.int32)
        self._work = {}
        self.newton_ncalls = 0
        self.lin_ncalls = 0
        self.work_counters["newton"] = DeviceWorkCounter(self._counters[0:1])
        self.work_counters["rhs"] = DeviceWorkCounter(self._counters[2:3])
        self.work_counters["linear"] = DeviceWorkCounter(self._counters[1:2])

    @property
    def ndim(self):
        return 2

    # the reference's plain-int totals (AllenCahn_2D_FD.py:127-130,202-203,347-348); the counts live on the device, so
    # reading one synchronises
    @property
    def newton_itercount(self):
        return int(self._counters[6].item())

    @property
    def lin_itercount(self):
        return int(self._counters[7].item())

    @classmethod
    def get_default_sweeper_class(cls):
        from .sweepers import generic_implicit

        return generic_implicit

    def eval_f_batch(self, us, ts, fs):
        self._be.allencahn_eval_f(self._lay, self.a_diag, self.a_off, 1.0 / self.eps**2, int(self.nu),
                                  [u.flat for u in us], [f.flat for f in fs])
        for _ in us:
            self.work_counters["rhs"]()

    def eval_f(self, u, t):
        """AllenCahn_2D_FD.py:207-228."""
        f = self.dtype_f(self.init)
        self.eval_f_batch([u], [t], [f])
        return f

    def solve_system_batch(self, rhs, factors, xs, ts=None):
        """Newton + inner CG for all given systems in ONE persistent launch, in place on xs
        (AllenCahn_2D_FD.py:137-205; several systems = the independent node solves of a diagonal QDelta)."""
        B = len(xs)
        if B not in self._work:
            self._work[B] = self._be.newton_workspace(self._lay, B)
        log = getattr(self, "solve_log", None)  # bench.py: per-launch device timing + iteration counts
        counters = self._counters[4:6]
        counters.zero_()
        if log is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        self._newton_launch(list(factors), rhs, xs, self._work[B], counters)
        if log is not None:
            ev1.record()
            log.append((ev0, ev1, counters.clone()))
        self._count_newton(counters)
        self.newton_ncalls += B

    def _newton_launch(self, factors, rhs, xs, work, counters):
        self._be.allencahn_newton_solve(self._lay, factors, self.a_diag, self.a_off, 1.0 / self.eps**2,
                                        int(self.nu), [r.flat for r in rhs], [x.flat for x in xs], self.newton_tol,
                                        self.newton_maxiter, self.lin_tol, self.lin_maxiter, self.inexact_linear_ratio,
                                        work, counters)

    def _count_newton(self, counters):
        self._counters[0:2] += counters  # work_counters['newton'], ['linear'] (AllenCahn_2D_FD.py:188,194)
        self._counters[6:7] += counters[0:1]  # newton_itercount (:203)

    def solve_system(self, rhs, factor, u0, t):
        me = self.dtype_u(u0)
        self.solve_system_batch([rhs], [factor], [me], [t])
        return me

    def _rhs_total(self, f):
        """Full right-hand side as one field (the semi-implicit class adds its two components)."""
        return f

    def u_exact(self, t, u_init=None, t_init=None):
        """AllenCahn_2D_FD.py:230-257: the tanh circle at t = 0; for t > 0 a reference solution from scipy's
        ``solve_ivp`` with tolerances of 100 ulp (core/problem.py:118-152), its right-hand side evaluated by the device
        ``eval_f`` (each call uploads the state and downloads f: a checking aid, as in the reference, not a fast path)."""
        me = self.dtype_u(self.init, val=0.0)
        if t > 0:
            from scipy.integrate import solve_ivp

            rhs_counter = self.work_counters["rhs"]

            def eval_rhs(tt, u):
                v = self.dtype_u(self.init)
                v[:] = u.reshape(self.nvars)
                out = self._rhs_total(self.eval_f(v, tt)).get().flatten()
                rhs_counter.decrement()  # reference evaluations are not part of the run's work
                return out

            u0 = self.u_exact(0.0).get() if u_init is None else np.asarray(u_init.get() if hasattr(u_init, "get") else u_init) * 1.0
            tol = 100 * np.finfo(float).eps
            sol = solve_ivp(eval_rhs, (0 if t_init is None else t_init, t), u0.flatten(), atol=tol, rtol=tol)
            me[:] = sol.y[:, -1].reshape(self.nvars)
            return me
        X, Y = np.meshgrid(self.xvalues, self.xvalues)
        me[:] = np.tanh((self.radius - np.sqrt(X**2 + Y**2)) / (np.sqrt(2) * self.eps))
        return me


class AllenCahnSemiMixin(AllenCahnMixin):
    """``allencahn_semiimplicit`` (AllenCahn_2D_FD.py:261-376): the Laplacian implicit - a periodic CG solve of
    (I - factor A) u = rhs with ``lin_tol`` / ``lin_maxiter`` on the TMA-pipelined solver -, the reaction term explicit."""

    dtype_f = imex_mesh

    @classmethod
    def get_default_sweeper_class(cls):
        from .sweepers import imex_1st_order

        return imex_1st_order

    _fcomps = ("impl", "expl")
    _count_rhs = True

    def _rhs_total(self, f):
        return getattr(f, self._fcomps[0]) + getattr(f, self._fcomps[1])  # AllenCahn_2D_FD.py:364-366

    def eval_f_batch(self, us, ts, fs):
        c1, c2 = self._fcomps
        self._be.allencahn_eval_f(self._lay, self.a_diag, self.a_off, 1.0 / self.eps**2, int(self.nu),
                                  [u.flat for u in us], [getattr(f, c1).flat for f in fs], [getattr(f, c2).flat for f in fs])
        if self._count_rhs:
            for _ in us:
                self.work_counters["rhs"]()

    def solve_system_batch(self, rhs, factors, xs, ts=None):
        B = len(xs)
        key = ("cg", B)
        if key not in self._work:
            self._work[key] = self._be.cg_workspace(self._lay, B)
            self._cg_counters = self._be.zeros(8, dtype=torch.int32)
        counters = self._cg_counters[:B]
        counters.zero_()
        log = getattr(self, "solve_log", None)
        if log is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        self._be.heat_cg_solve(self._lay, BC_CODES["periodic"], [1.0 - f * self.a_diag for f in factors],
                               [-(f * self.a_off) for f in factors], [r.flat for r in rhs], [x.flat for x in xs],
                               self.lin_tol, self.lin_maxiter, self._work[key], counters)
        if log is not None:
            ev1.record()
            log.append((ev0, ev1, counters.clone()))
        self._count_linear(counters.sum(dtype=torch.int32))
        self.lin_ncalls += B

    def _count_linear(self, total):
        self._counters[1:2] += total  # work_counters['linear'] (AllenCahn_2D_FD.py:327-331)
        self._counters[7:8] += total  # lin_itercount (:348)


class AllenCahnSemiV2Mixin(AllenCahnMixin):
    """``allencahn_semiimplicit_v2`` (AllenCahn_2D_FD.py:380-484): ``Delta u - u^(nu+1)/eps^2`` implicit - the batched
    Newton kernel with the Jacobian of that part; the inner CG runs to ``lin_tol`` with scipy's default iteration cap of
    ten times the system size and the reference counts neither its iterations nor the Newton steps in ``work_counters``
    (only ``newton_itercount`` / ``newton_ncalls``, :481-482) -, ``u/eps^2`` explicit."""

    dtype_f = imex_mesh

    @classmethod
    def get_default_sweeper_class(cls):
        from .sweepers import imex_1st_order

        return imex_1st_order

    _fcomps = ("impl", "expl")

    def _rhs_total(self, f):
        return getattr(f, self._fcomps[0]) + getattr(f, self._fcomps[1])

    def eval_f_batch(self, us, ts, fs):  # :402-424 (no work counter there)
        c1, c2 = self._fcomps
        self._be.allencahn_eval_f(self._lay, self.a_diag, self.a_off, 1.0 / self.eps**2, int(self.nu),
                                  [u.flat for u in us], [getattr(f, c1).flat for f in fs], [getattr(f, c2).flat for f in fs],
                                  split=2)

    def _newton_launch(self, factors, rhs, xs, work, counters):
        n = self.nvars[0] * self.nvars[1]
        self._be.allencahn_newton_solve(self._lay, factors, self.a_diag, self.a_off, 1.0 / self.eps**2,
                                        int(self.nu), [r.flat for r in rhs], [x.flat for x in xs], self.newton_tol,
                                        self.newton_maxiter, self.lin_tol, min(10 * n, 2**31 - 1), None, work, counters,
                                        variant=1)

    def _count_newton(self, counters):
        self._counters[6:7] += counters[0:1]


class _TwoSolves:
    """``solve_system_1`` / ``solve_system_2`` of the multi-implicit problem classes on top of their in-place batch forms."""

    dtype_f = comp2_mesh
    _fcomps = ("comp1", "comp2")

    @classmethod
    def get_default_sweeper_class(cls):
        from .sweepers import multi_implicit

        return multi_implicit

    def solve_system_1(self, rhs, factor, u0, t):
        me = self.dtype_u(u0)
        self.solve_system_1_batch([rhs], [factor], [me], [t])
        return me

    def solve_system_2(self, rhs, factor, u0, t):
        me = self.dtype_u(u0)
        self.solve_system_2_batch([rhs], [factor], [me], [t])
        return me

    def solve_system(self, rhs, factor, u0, t):
        raise ProblemError(f"{type(self).__name__} splits the right-hand side in two implicit parts: use solve_system_1 / "
                           "solve_system_2 (multi_implicit sweeper)")

    def solve_system_batch(self, rhs, factors, xs, ts=None):
        self.solve_system(rhs, factors, xs, ts)  # same error for the batched entry of the single-solve sweepers


class AllenCahnMultiMixin(_TwoSolves, AllenCahnSemiMixin):
    """``allencahn_multiimplicit`` (AllenCahn_2D_FD.py:487-651): both parts of the right-hand side implicit, one after
    the other - ``solve_system_1``: the periodic Laplacian on the TMA-pipelined CG (``lin_tol`` / ``lin_maxiter``, counted
    in ``lin_itercount`` only, :585-590); ``solve_system_2``: the reaction term by a point-wise Newton iteration with the
    reference's global stopping rule in one persistent launch (``sdcb200_allencahn_reaction_newton``; counted in
    ``newton_itercount`` only, :648-649).  ``eval_f`` counts nothing (:510-532)."""

    _count_rhs = False

    def solve_system_1_batch(self, rhs, factors, xs, ts=None):
        AllenCahnSemiMixin.solve_system_batch(self, rhs, factors, xs, ts)

    def _count_linear(self, total):
        self._counters[7:8] += total

    def solve_system_2_batch(self, rhs, factors, xs, ts=None):
        if "react" not in self._work:
            self._work["react"] = self._be.reaction_workspace()
        self._be.allencahn_reaction_newton(list(factors), 1.0 / self.eps**2, int(self.nu), [r.vol for r in rhs],
                                           [x.vol for x in xs], self.newton_tol, self.newton_maxiter, self._work["react"],
                                           self._counters[6:7])
        self.newton_ncalls += len(xs)


class AllenCahnMultiV2Mixin(_TwoSolves, AllenCahnSemiV2Mixin):
    """``allencahn_multiimplicit_v2`` (AllenCahn_2D_FD.py:655-776): ``solve_system_1`` is the Newton system of
    ``allencahn_semiimplicit_v2`` (batched Newton kernel, ``variant = 1``); ``solve_system_2`` is the closed-form
    ``rhs / (1 - factor / eps^2)`` (:774)."""

    def solve_system_1_batch(self, rhs, factors, xs, ts=None):
        AllenCahnMixin.solve_system_batch(self, rhs, factors, xs, ts)

    def solve_system_2_batch(self, rhs, factors, xs, ts=None):
        for r, factor, x in zip(rhs, factors, xs):
            self._be.axpby(1.0 / (1.0 - factor * 1.0 / self.eps**2), r.vol, 0.0, None, x.vol)
            x._touch()


def _bind(base):
    """Concrete problem classes over a given ``Problem`` base class."""
    ns = {}
    for name, mixin in (("heatNd_unforced", HeatMixin), ("heatNd_forced", HeatForcedMixin),
                        ("advectionNd", AdvectionMixin),
                        ("allencahn_fullyimplicit", AllenCahnMixin), ("allencahn_semiimplicit", AllenCahnSemiMixin),
                        ("allencahn_semiimplicit_v2", AllenCahnSemiV2Mixin),
                        ("allencahn_multiimplicit", AllenCahnMultiMixin),
                        ("allencahn_multiimplicit_v2", AllenCahnMultiV2Mixin)):
        ns[name] = type(name, (mixin, base), {"__doc__": mixin.__doc__, "__module__": __name__})
    return ns


from .core import Problem as _Problem  # noqa: E402

globals().update(_bind(_Problem))
