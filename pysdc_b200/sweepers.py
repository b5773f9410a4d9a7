"""SDC sweepers ``generic_implicit`` and ``imex_1st_order`` on the device.

Same class names, constructor, attributes (``coll``, ``QI``, ``QE``, ``params``) and methods as the reference
(``pySDC/implementations/sweeper_classes/generic_implicit.py``, ``imex_1st_order.py``, base ``pySDC/core/sweeper.py``):
``predict``, ``integrate``, ``update_nodes``, ``compute_residual``, ``compute_end_point``.  They read and write only
``L.u, L.f, L.tau, L.uend, L.residual, L.status.{residual, updated, unlocked}`` like the originals, so the reference's
controllers, transfer classes and convergence controllers can drive them unchanged.

What changes is how the work is done:

* all "known terms" of a sweep — ``u0 + dt (Q - QDelta) F(u^k) + tau`` for every node (generic_implicit.py:70-82) — are
  one launch of the fused collocation kernel (each of the M*C right-hand sides read once, M results written once)
  instead of ~2M^2 axpy passes with temporaries;
* when QDelta is diagonal (MIN-SR-NS, MIN-SR-FLEX, IEpar, ...; ``sweeper.parallelizable``, core/sweeper.py:108-109) the M
  node systems are independent: they are solved by ONE persistent batched-CG launch and their right-hand sides are
  re-evaluated by one stencil launch; lower-triangular QDelta (LU, IE) keeps the sequential node order of
  generic_implicit.py:85-98 with one solve launch per node;
* the residual (core/sweeper.py:164-215) is one fused pass producing the M max-norms on the device; the single
  device->host read of a sweep happens here, because ``L.status.residual`` has to be a Python float for
  ``CheckConvergence`` (convergence_controller_classes/check_convergence.py:75-76).

Rounding: the node combinations go through ``sdcb200_colloc_sweep``, which performs the reference's floating-point
operations in the reference's order (scalar coefficient first, every product and sum rounded on its own, quadrature
terms before QDelta terms before ``u[0]`` before ``tau``), so given identical solves a sweep reproduces numpy bit for bit.

Solutions and right-hand sides are updated in place in the buffers ``L.u[m]`` / ``L.f[m]`` already own — unless somebody
else holds a reference to such a field (the reference REBINDS ``L.u[m+1]`` to a fresh object in every sweep, so code like
``uold[1:] = L.u[1:]`` in controller_MPI.py:475 keeps the old values): then the sweep switches to a fresh buffer first.
"""
import sys

import numpy as np

from .backend import get_backend
from .errors import ParameterError


class _LazyResiduals:
    """``L.residual`` as the reference leaves it after ``compute_residual`` (core/sweeper.py:188-193: the M residual
    fields), materialised only when somebody reads it: the sweep itself needs the max-norms only, and writing M more
    fields per sweep would add a third to its streaming traffic."""

    def __init__(self, sweeper):
        self._sweeper, self._fields = sweeper, None

    def _materialise(self):
        if self._fields is None:
            self._fields = self._sweeper._residual_fields()
        return self._fields

    def __getitem__(self, m):
        return self._materialise()[m]

    def __len__(self):
        return self._sweeper.coll.num_nodes

    def __iter__(self):
        return iter(self._materialise())


class _SweepCommon:
    """Methods shared by both sweepers; ``self.coll / self.params / self.level / self.QI`` come from the base class."""

    imex = False
    _comps = ("impl", "expl")  # names of the two right-hand-side components when imex

    @property
    def _ncomp(self):
        return 2 if self.imex else 1

    # ---- helpers ----------------------------------------------------------------------------------------------------
    def _f_inputs(self, L, first=1):
        """Flat device views of f[first..M], node-major then component (impl, expl)."""
        ins = []
        for j in range(first, self.coll.num_nodes + 1):
            f = L.f[j]
            if self.imex:
                ins += [getattr(f, self._comps[0]).flat, getattr(f, self._comps[1]).flat]
            else:
                ins.append(f.flat)
        return ins

    @staticmethod
    def _own(lst, i):
        """``lst[i]`` ready to be overwritten in place: if anything besides the level's list refers to the object (an
        ``uold`` list, a hook's record, a user variable) it is replaced by a fresh copy first, which leaves the other
        holder with the old values exactly as the reference's rebinding does."""
        x = lst[i]
        if sys.getrefcount(x) > 3:  # the list, the local name, getrefcount's argument
            lst[i] = type(x)(x)
        del x
        return lst[i]

    @staticmethod
    def _field_key(x):
        return None if x is None else (id(x), x._buf._version, getattr(x, "_kver", 0))

    def _residual_key(self, L):
        """Identity + version of everything compute_residual reads (core/sweeper.py:164-215)."""
        return (L.dt, L.params.residual_type, tuple(self._field_key(u) for u in L.u),
                tuple(self._field_key(f) for f in L.f[1:]), tuple(self._field_key(t) for t in L.tau),
                self.coll.Qmat.tobytes())

    def _scratch(self, L, count):
        """M reusable right-hand-side fields per level (never visible to the caller)."""
        key = (id(L), L.prob.init[0])
        pool = self.__dict__.setdefault("_rhs_pool", {})
        if key not in pool or len(pool[key]) < count:
            pool[key] = [L.prob.dtype_u(L.prob.init) for _ in range(count)]
        return pool[key]

    # ---- predictor (core/sweeper.py:125-162) ------------------------------------------------------------------------
    def predict(self):
        L = self.level
        P = L.prob
        M = self.coll.num_nodes
        guess = self.params.initial_guess
        times = [L.time] + [L.time + L.dt * self.coll.nodes[m] for m in range(M)]
        if guess == "spread":
            for m in range(1, M + 1):
                L.u[m] = P.dtype_u(L.u[0])
            if hasattr(P, "eval_f_batch"):
                for m in range(M + 1):
                    L.f[m] = P.dtype_f(P.init)
                P.eval_f_batch(L.u, times, L.f)
            else:
                for m in range(M + 1):
                    L.f[m] = P.eval_f(L.u[m], times[m])
        elif guess in ("copy", "zero", "random"):
            L.f[0] = P.eval_f(L.u[0], L.time)
            for m in range(1, M + 1):
                if guess == "copy":
                    L.u[m] = P.dtype_u(L.u[0])
                    L.f[m] = P.dtype_f(L.f[0])
                elif guess == "zero":
                    L.u[m] = P.dtype_u(init=P.init, val=0.0)
                    L.f[m] = P.dtype_f(init=P.init, val=0.0)
                else:
                    L.u[m] = P.dtype_u(init=P.init, val=self.rng.rand(1)[0])
                    L.f[m] = P.dtype_f(init=P.init, val=self.rng.rand(1)[0])
        else:
            raise ParameterError(f"initial_guess option {guess} not implemented")
        L.status.unlocked = True
        L.status.updated = True
        self._res_cache = None

    # ---- integrate (generic_implicit.py:29-49, imex_1st_order.py:37-55) ---------------------------------------------
    def integrate(self):
        L = self.level
        P = L.prob
        M = self.coll.num_nodes
        me = [P.dtype_u(P.init) for _ in range(M)]
        get_backend().colloc_sweep(self._f_inputs(L), self._ncomp, [x.flat for x in me], Wq=L.dt * self.coll.Qmat[1:, 1:])
        return me

    # ---- one sweep (generic_implicit.py:51-103, imex_1st_order.py:57-108) -------------------------------------------
    def update_nodes(self):
        L = self.level
        P = L.prob
        assert L.status.unlocked
        be = get_backend()
        M = self.coll.num_nodes
        dt = L.dt
        Q, QI = self.coll.Qmat, self.QI
        QE = self.QE if self.imex else None
        times = [L.time + dt * self.coll.nodes[m] for m in range(M)]
        alphas = [dt * QI[m + 1, m + 1] for m in range(M)]
        batched = hasattr(P, "solve_system_batch") and hasattr(P, "eval_f_batch")

        # known terms of every node: u0 + dt*(Q - QDelta) F(u^k) + tau, one fused pass
        # (integrate(), then `integral[m] -= dt*QDelta[m+1, j] f[j]` for all j, `+= u[0]`, `+= tau[m]`: generic_implicit.py:
        # 70-82, imex_1st_order.py:77-88, in that order of operations)
        rhs = self._scratch(L, M)
        taus = [None if t is None else t.flat for t in L.tau]
        if self.imex:
            qd = dict(Wi=-QI[1:, 1:], We=-QE[1:, 1:], dt2=dt)
        else:
            qd = dict(Wi=-(dt * QI[1:, 1:]))
        be.colloc_sweep(self._f_inputs(L), self._ncomp, [r.flat for r in rhs], Wq=dt * Q[1:, 1:], base=L.u[0].flat,
                        adds=taus if any(t is not None for t in taus) else None, **qd)
        for m in range(1, M + 1):
            self._own(L.u, m)
            self._own(L.f, m)

        strictly_lower_empty = not np.any(np.tril(QI[1:, 1:], k=-1)) and (QE is None or not np.any(np.tril(QE[1:, 1:], k=-1)))
        if batched and strictly_lower_empty and (self.imex or all(a != 0 for a in alphas)):
            # diagonal QDelta: the M node systems are independent -> one batched solve, one batched f evaluation
            us, fs = L.u[1:], L.f[1:]
            P.solve_system_batch(rhs[:M], alphas, us, times)
            P.eval_f_batch(us, times, fs)
        else:
            for m in range(M):
                if m > 0:
                    # add dt*QDelta[m+1, j] f(u_j^{k+1}) for the nodes j <= m already updated in this sweep
                    # (generic_implicit.py:87-89, imex_1st_order.py:92-95), in place on rhs[m]
                    ins = self._f_inputs(L)[: m * self._ncomp]
                    if self.imex:
                        qd = dict(Wi=QI[m + 1: m + 2, 1: m + 1], We=QE[m + 1: m + 2, 1: m + 1], dt2=dt)
                    else:
                        qd = dict(Wi=dt * QI[m + 1: m + 2, 1: m + 1])
                    if np.any(qd["Wi"]) or (self.imex and np.any(qd["We"])):
                        be.colloc_sweep(ins, self._ncomp, [rhs[m].flat], base=rhs[m].flat, base_first=True, **qd)
                if alphas[m] == 0 and not self.imex:
                    L.u[m + 1][:] = rhs[m]  # generic_implicit.py:93-94
                elif batched:
                    P.solve_system_batch([rhs[m]], [alphas[m]], [L.u[m + 1]], [times[m]])
                else:
                    L.u[m + 1] = P.solve_system(rhs[m], alphas[m], L.u[m + 1], times[m])
                if batched:
                    P.eval_f_batch([L.u[m + 1]], [times[m]], [L.f[m + 1]])
                else:
                    L.f[m + 1] = P.eval_f(L.u[m + 1], times[m])
        L.status.updated = True
        self._res_cache = None
        return None

    # ---- residual (core/sweeper.py:164-215) -------------------------------------------------------------------------
    def compute_residual(self, stage=""):
        L = self.level
        if stage in self.params.skip_residual_computation:
            L.status.residual = 0.0 if L.status.residual is None else L.status.residual
            return None
        # the reference's controllers ask for the residual twice per iteration with nothing changed in between
        # (controller_nonMPI.py:493 after :573): the second request is answered from the first when neither a sweep
        # (L.status.updated) nor anybody else touched the fields the residual reads
        key = self._residual_key(L)
        cache = getattr(self, "_res_cache", None)
        if (not L.status.updated and cache is not None and cache[0] == key
                and not getattr(self.params, "store_residual", False)):
            L.status.residual = cache[1]
            return None
        be = get_backend()
        M = self.coll.num_nodes
        Wq = L.dt * self.coll.Qmat[1:, 1:]
        taus = [None if t is None else t.flat for t in L.tau]
        res_out = None
        if getattr(self.params, "store_residual", False):
            L.residual = [L.prob.dtype_u(L.prob.init) for _ in range(M)]
            res_out = [r.flat for r in L.residual]
        else:
            L.residual = _LazyResiduals(self)
        if "_resnorm" not in self.__dict__:
            self._resnorm = be.zeros(9)
        norms_dev = self._resnorm
        be.colloc_residual(Wq, self._f_inputs(L), self._ncomp, L.u[0].flat, [u.flat for u in L.u[1:]],
                           taus if any(t is not None for t in taus) else None, res_out, norms_dev[:M])
        rtype = L.params.residual_type
        if rtype.endswith("_rel"):
            be.maxabs_async(L.u[0].vol, norms_dev[8:9])
        comm = L.u[0].comm
        distributed = comm is not None and getattr(comm, "size", 1) > 1
        if distributed and hasattr(comm, "allreduce_device"):
            from .comm import MAX
            comm.allreduce_device(norms_dev, MAX)  # one MAX all-reduce on the device vector, then the single read
            distributed = False
        host = norms_dev.cpu().tolist()  # the one device->host read of a sweep
        res_norm, u0_norm = host[:M], host[8]
        if distributed:
            from .comm import MAX
            res_norm = comm.allreduce(res_norm, op=MAX)
            u0_norm = comm.allreduce(u0_norm, op=MAX)
        if rtype == "full_abs":
            L.status.residual = max(res_norm)
        elif rtype == "last_abs":
            L.status.residual = res_norm[-1]
        elif rtype == "full_rel":
            L.status.residual = max(res_norm) / u0_norm
        elif rtype == "last_rel":
            L.status.residual = res_norm[-1] / u0_norm
        else:
            raise ParameterError(f"residual_type = {rtype} not implemented, choose full_abs, last_abs, full_rel or "
                                 "last_rel instead")
        L.status.updated = False
        self._res_cache = (key, L.status.residual)
        return None

    def _residual_fields(self):
        """The M residual fields of the level's current state (what the reference stores in ``L.residual``)."""
        L = self.level
        be = get_backend()
        M = self.coll.num_nodes
        taus = [None if t is None else t.flat for t in L.tau]
        out = [L.prob.dtype_u(L.prob.init) for _ in range(M)]
        be.colloc_residual(L.dt * self.coll.Qmat[1:, 1:], self._f_inputs(L), self._ncomp, L.u[0].flat,
                           [u.flat for u in L.u[1:]], taus if any(t is not None for t in taus) else None,
                           [r.flat for r in out], be.zeros(M))
        return out

    # ---- end point (generic_implicit.py:105-131, imex_1st_order.py:110-137) -----------------------------------------
    def compute_end_point(self):
        L = self.level
        P = L.prob
        if self.coll.right_is_node and not self.params.do_coll_update:
            L.uend = P.dtype_u(L.u[-1])
        else:
            # uend = u[0]; uend += dt*w_m f[m+1] for all m; uend += tau[-1]   (generic_implicit.py:123-129)
            L.uend = P.dtype_u(P.init)
            tau = None if L.tau[-1] is None else [L.tau[-1].flat]
            get_backend().colloc_sweep(self._f_inputs(L), self._ncomp, [L.uend.flat],
                                       Wq=(L.dt * self.coll.weights)[None, :], base=L.u[0].flat, adds=tau, base_first=True)
        return None


class GenericImplicitMixin(_SweepCommon):
    """generic_implicit.py:4-27."""

    imex = False

    def __init__(self, params, level):
        if "QI" not in params:
            params["QI"] = "IE"
        super().__init__(params, level)
        self.QI = self.get_Qdelta_implicit(qd_type=self.params.QI)


class Imex1stOrderMixin(_SweepCommon):
    """imex_1st_order.py:6-35."""

    imex = True

    def __init__(self, params, level):
        if "QI" not in params:
            params["QI"] = "IE"
        if "QE" not in params:
            params["QE"] = "EE"
        super().__init__(params, level)
        self.QI = self.get_Qdelta_implicit(qd_type=self.params.QI)
        self.QE = self.get_Qdelta_explicit(qd_type=self.params.QE)


class MultiImplicitMixin(_SweepCommon):
    """``multi_implicit`` (sweeper_classes/multi_implicit.py:4-160): first-order sweeper for a right-hand side split in two
    parts ``comp1`` / ``comp2`` that are BOTH treated implicitly, one after the other, with their own QDelta (``Q1``,
    ``Q2``) and their own solver (``P.solve_system_1`` / ``P.solve_system_2``).  ``integrate``, the residual, the
    predictor and the end point are the two-component forms of the common sweeper (``f.comp1 + f.comp2``)."""

    imex = True
    _comps = ("comp1", "comp2")

    def __init__(self, params, level):
        if "Q1" not in params:
            params["Q1"] = "IE"
        if "Q2" not in params:
            params["Q2"] = "IE"
        super().__init__(params, level)
        self.Q1 = self.get_Qdelta_implicit(qd_type=self.params.Q1)
        self.Q2 = self.get_Qdelta_implicit(qd_type=self.params.Q2)

    def update_nodes(self):
        """multi_implicit.py:60-132.  Known terms ``u0 + dt (Q F - Q1 F1) + tau`` of all nodes and ``dt Q2 F2`` of all
        nodes: two fused launches; per node one small launch per implicit part for the terms of the nodes already updated
        (:108-110,:118-120), in place on the scratch right-hand sides."""
        L = self.level
        P = L.prob
        assert L.status.unlocked
        be = get_backend()
        M = self.coll.num_nodes
        dt = L.dt
        Q, Q1, Q2 = self.coll.Qmat, self.Q1, self.Q2
        scratch = self._scratch(L, 2 * M)
        integral, q2int = scratch[:M], scratch[M: 2 * M]
        taus = [None if t is None else t.flat for t in L.tau]
        be.colloc_sweep(self._f_inputs(L), 2, [r.flat for r in integral], Wq=dt * Q[1:, 1:], Wi=-Q1[1:, 1:],
                        We=np.zeros((M, M)), dt2=dt, base=L.u[0].flat,
                        adds=taus if any(t is not None for t in taus) else None)
        be.colloc_sweep([L.f[j].comp2.flat for j in range(1, M + 1)], 1, [q.flat for q in q2int], Wi=dt * Q2[1:, 1:])
        batched = hasattr(P, "solve_system_1_batch") and hasattr(P, "solve_system_2_batch") and hasattr(P, "eval_f_batch")
        for m in range(1, M + 1):
            self._own(L.u, m)
            self._own(L.f, m)
        for m in range(M):
            t = L.time + dt * self.coll.nodes[m]
            if m > 0 and np.any(Q1[m + 1, 1: m + 1]):
                be.colloc_sweep([L.f[j].comp1.flat for j in range(1, m + 1)], 1, [integral[m].flat],
                                Wi=dt * Q1[m + 1: m + 2, 1: m + 1], base=integral[m].flat, base_first=True)
            if batched:
                P.solve_system_1_batch([integral[m]], [dt * Q1[m + 1, m + 1]], [L.u[m + 1]], [t])
            else:
                L.u[m + 1] = P.solve_system_1(integral[m], dt * Q1[m + 1, m + 1], L.u[m + 1], t)
            # rhs = u - Q2int[m] + sum_{j<=m} dt*Q2[m+1, j] f[j].comp2   (:116-120), in place on q2int[m]
            ins = [q2int[m].flat] + [L.f[j].comp2.flat for j in range(1, m + 1)]
            W = np.concatenate([[-1.0], dt * Q2[m + 1, 1: m + 1]])[None, :]
            be.colloc_sweep(ins, 1, [q2int[m].flat], Wi=W, base=L.u[m + 1].flat, base_first=True)
            if batched:
                P.solve_system_2_batch([q2int[m]], [dt * Q2[m + 1, m + 1]], [L.u[m + 1]], [t])
                P.eval_f_batch([L.u[m + 1]], [t], [L.f[m + 1]])
            else:
                L.u[m + 1] = P.solve_system_2(q2int[m], dt * Q2[m + 1, m + 1], L.u[m + 1], t)
                L.f[m + 1] = P.eval_f(L.u[m + 1], t)
        L.status.updated = True
        self._res_cache = None
        return None


class NodeParallelMixin:
    """``SweeperMPI`` (sweeper_classes/generic_implicit_MPI.py:8-164): parallel across the method, one collocation node
    per rank / GPU, for diagonal QDelta.  Placed in front of the sweeper mix-in like the reference's multiple inheritance
    (``class generic_implicit_MPI(SweeperMPI, generic_implicit)``, :167).

    The reference moves every node's right-hand side M times per quadrature (M ``comm.Reduce`` of full vectors,
    :176-196 — and twice per sweep, for ``update_nodes`` and for the residual).  Here the right-hand side of the own
    node is ALL-GATHERED once after it changed (one NCCL collective over NVLink, each field crosses once), after which
    the quadrature of the own node, the residual and the end-point update are local launches of the fused collocation
    kernels on the gathered fields.  The sums run over the nodes in the serial sweeper's order j = 1..M (an MPI reduction
    has no defined order), so a node-parallel run reproduces the serial diagonal-QDelta sweep bit for bit.  Like in the
    reference only ``L.u[rank+1]`` / ``L.f[rank+1]`` (and index 0) exist on a rank."""

    def __init__(self, params, level):
        if "comm" not in params:  # generic_implicit_MPI.py:35-37: MPI.COMM_WORLD - here the mpi4py-style communicator
            # over all ranks of the torch.distributed job (other reference code, e.g. base_transfer_MPI, talks to it
            # through the mpi4py surface: Reduce, Bcast, ...)
            import sys

            if "mpi4py.MPI" in sys.modules and hasattr(getattr(sys.modules["mpi4py.MPI"], "COMM_WORLD", None), "tc"):
                params["comm"] = sys.modules["mpi4py.MPI"].COMM_WORLD
            else:
                from .mpi_facade.mpi4py import MPI

                params["comm"] = MPI.COMM_WORLD
        super().__init__(params, level)
        if self.params.comm.size != self.coll.num_nodes:
            raise NotImplementedError(
                f"The communicator in the {type(self).__name__} sweeper needs to have one rank for each node as of now! "
                f"That means we need {self.coll.num_nodes} nodes, but got {self.params.comm.size} processes.")
        self._fall, self._fall_key = None, None

    @property
    def comm(self):
        return self.params.comm

    @property
    def rank(self):
        return self.comm.rank

    @property
    def _tc(self):
        return getattr(self.comm, "tc", self.comm)  # the mpi4py facade wraps a TorchComm

    # ---- the gathered right-hand sides ------------------------------------------------------------------------------
    def _all_f(self, L):
        """f of every node, current: one all-gather whenever the own node's f changed since the last one (every rank
        runs the same program, so the decision is the same on all of them)."""
        own = L.f[self.rank + 1]
        key = self._field_key(own)
        if self._fall is None:
            self._fall = [L.prob.dtype_f(L.prob.init) for _ in range(self.coll.num_nodes)]
        if self._fall_key != key:
            self._tc.Allgather_fields(own, self._fall)
            self._fall_key = key
        return self._fall

    def _gathered_inputs(self, L):
        ins = []
        for f in self._all_f(L):
            ins += [getattr(f, self._comps[0]).flat, getattr(f, self._comps[1]).flat] if self.imex else [f.flat]
        return ins

    # ---- predictor (generic_implicit_MPI.py:124-157) ----------------------------------------------------------------
    def predict(self):
        L = self.level
        P = L.prob
        m = self.rank
        guess = self.params.initial_guess
        t_m = L.time + L.dt * self.coll.nodes[m]
        if guess == "spread":
            L.u[m + 1] = P.dtype_u(L.u[0])
            if hasattr(P, "eval_f_batch"):
                L.f[0], L.f[m + 1] = P.dtype_f(P.init), P.dtype_f(P.init)
                P.eval_f_batch([L.u[0], L.u[m + 1]], [L.time, t_m], [L.f[0], L.f[m + 1]])
            else:
                L.f[0] = P.eval_f(L.u[0], L.time)
                L.f[m + 1] = P.eval_f(L.u[m + 1], t_m)
        elif guess == "copy":
            L.f[0] = P.eval_f(L.u[0], L.time)
            L.u[m + 1] = P.dtype_u(L.u[0])
            L.f[m + 1] = P.dtype_f(L.f[0])
        elif guess == "zero":
            L.f[0] = P.eval_f(L.u[0], L.time)
            L.u[m + 1] = P.dtype_u(init=P.init, val=0.0)
            L.f[m + 1] = P.dtype_f(init=P.init, val=0.0)
        else:
            raise ParameterError(f"initial_guess option {guess} not implemented")
        L.status.unlocked = True
        L.status.updated = True
        self._res_cache, self._fall_key = None, None

    # ---- integrate (generic_implicit_MPI.py:176-196, imex_1st_order_MPI.py:14-38) -----------------------------------
    def integrate(self, last_only=False):
        """The own node's row of ``dt Q F`` (ONE field, not a list); with ``last_only`` only the last rank gets it."""
        L = self.level
        me = L.prob.dtype_u(L.prob.init, val=0.0)
        ins = self._gathered_inputs(L)
        if not last_only or self.rank == self.coll.num_nodes - 1:
            get_backend().colloc_sweep(ins, self._ncomp, [me.flat],
                                       Wq=L.dt * self.coll.Qmat[self.rank + 1: self.rank + 2, 1:])
        return me

    # ---- one sweep (generic_implicit_MPI.py:198-239, imex_1st_order_MPI.py:40-88) -----------------------------------
    def update_nodes(self):
        L = self.level
        P = L.prob
        assert L.status.unlocked
        r = self.rank
        dt = L.dt
        alpha = dt * self.QI[r + 1, r + 1]
        t_r = L.time + dt * self.coll.nodes[r]
        # rhs = (dt Q F)[r] - dt*QDelta[r+1, r+1] f[r+1] + u[0] + tau[r], one fused pass over the gathered fields
        M = self.coll.num_nodes
        Wi = np.zeros((1, M))
        Wi[0, r] = -self.QI[r + 1, r + 1]
        if self.imex:
            qd = dict(Wi=Wi, We=np.zeros((1, M)), dt2=dt)  # QE = PIC (imex_1st_order_MPI.py:9-12)
        else:
            qd = dict(Wi=dt * Wi)
        rhs = self._scratch(L, 1)[0]
        tau = None if L.tau[r] is None else [L.tau[r].flat]
        get_backend().colloc_sweep(self._gathered_inputs(L), self._ncomp, [rhs.flat], Wq=dt * self.coll.Qmat[r + 1: r + 2, 1:],
                                   base=L.u[0].flat, adds=tau, **qd)
        self._own(L.u, r + 1)
        self._own(L.f, r + 1)
        if hasattr(P, "solve_system_batch") and hasattr(P, "eval_f_batch"):
            P.solve_system_batch([rhs], [alpha], [L.u[r + 1]], [t_r])
            P.eval_f_batch([L.u[r + 1]], [t_r], [L.f[r + 1]])
        else:
            L.u[r + 1] = P.solve_system(rhs, alpha, L.u[r + 1], t_r)
            L.f[r + 1] = P.eval_f(L.u[r + 1], t_r)
        L.status.updated = True
        self._res_cache, self._fall_key = None, None
        return None

    # ---- residual (generic_implicit_MPI.py:80-122) ------------------------------------------------------------------
    def compute_residual(self, stage=None):
        L = self.level
        if stage in self.params.skip_residual_computation:
            L.status.residual = 0.0 if L.status.residual is None else L.status.residual
            return None
        rtype = L.params.residual_type
        if rtype not in ("full_abs", "last_abs", "full_rel", "last_rel"):
            raise NotImplementedError(f'residual type "{rtype}" not implemented!')
        r, M = self.rank, self.coll.num_nodes
        key = (L.dt, rtype, self._field_key(L.u[0]), self._field_key(L.u[r + 1]), self._field_key(L.f[r + 1]),
               self._field_key(L.tau[r]))
        cache = getattr(self, "_res_cache", None)
        if not L.status.updated and cache is not None and cache[0] == key:
            L.status.residual = cache[1]  # second request of an iteration (controller_nonMPI.py:493 after :573)
            return None
        be = get_backend()
        if "_resnorm" not in self.__dict__:
            self._resnorm = be.zeros(M + 1)
        norms = self._resnorm
        norms.zero_()
        ins = self._gathered_inputs(L)
        if rtype.startswith("full") or r == M - 1:
            be.colloc_residual(L.dt * self.coll.Qmat[r + 1: r + 2, 1:], ins, self._ncomp, L.u[0].flat, [L.u[r + 1].flat],
                               None if L.tau[r] is None else [L.tau[r].flat], None, norms[r: r + 1])
        if rtype.endswith("_rel"):
            be.maxabs_async(L.u[0].vol, norms[M: M + 1])
        from .comm import MAX

        self._tc.allreduce_device(norms, MAX)  # every rank gets all M node norms: one collective, then one read
        space = L.u[0].comm  # nodes x slabs: the norms of a node are maxima over the slabs of its field as well
        if space is not None and getattr(space, "size", 1) > 1 and hasattr(space, "allreduce_device"):
            space.allreduce_device(norms, MAX)
        host = norms.cpu().tolist()
        res = max(host[:M]) if rtype.startswith("full") else host[M - 1]
        L.status.residual = res / host[M] if rtype.endswith("_rel") else res
        L.status.updated = False
        self._res_cache = (key, L.status.residual)
        return None

    # ---- end point (generic_implicit_MPI.py:53-78,241-267, imex_1st_order_MPI.py:90-124) ----------------------------
    def compute_end_point(self):
        L = self.level
        P = L.prob
        root = self.comm.Get_size() - 1
        if self.coll.right_is_node and not self.params.do_coll_update:
            L.uend = P.dtype_u(L.u[-1]) if self.rank == root else P.dtype_u(L.u[0])
            self._tc.Bcast(L.uend, root=root)
        else:
            # uend = sum_m dt*w_m f[m+1]; uend += u[0]; uend += tau[-1] (broadcast from the last rank)
            L.uend = P.dtype_u(P.init)
            tau = None
            if L.tau[self.rank] is not None:
                self.communicate_tau_correction_for_full_interval()
                tau = [L.tau[-1].flat]
            get_backend().colloc_sweep(self._gathered_inputs(L), self._ncomp, [L.uend.flat],
                                       Wq=(L.dt * self.coll.weights)[None, :], base=L.u[0].flat, adds=tau)
        return None

    def communicate_tau_correction_for_full_interval(self):
        L = self.level
        if self.rank < self.comm.size - 1:
            L.tau[-1] = L.prob.u_init
        self._tc.Bcast(L.tau[-1], root=self.comm.size - 1)


class Imex1stOrderNodeParallelMixin(NodeParallelMixin):
    def __init__(self, params, level):
        super().__init__(params, level)
        assert self.params.QE == "PIC", (f"Only Picard is implemented for explicit preconditioner so far in "
                                         f"{type(self).__name__}! You chose \"{self.params.QE}\"")


def _bind(base):
    ns = {}
    for name, mixin in (("generic_implicit", GenericImplicitMixin), ("imex_1st_order", Imex1stOrderMixin),
                        ("multi_implicit", MultiImplicitMixin)):
        ns[name] = type(name, (mixin, base), {"__doc__": mixin.__doc__, "__module__": __name__})
    for name, par, mixin in (("generic_implicit_MPI", NodeParallelMixin, GenericImplicitMixin),
                             ("imex_1st_order_MPI", Imex1stOrderNodeParallelMixin, Imex1stOrderMixin)):
        ns[name] = type(name, (par, mixin, base), {"__doc__": NodeParallelMixin.__doc__, "__module__": __name__})
    return ns


from .core import Sweeper as _Sweeper  # noqa: E402

globals().update(_bind(_Sweeper))
