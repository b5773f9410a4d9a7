"""Time-parallel controller: one time step per rank (one rank per GPU), SDC / MLSDC / MSSDC / PFASST.

Same constructor, ``run`` signature, stage sequence and message pattern as the reference's ``controller_MPI``
(``pySDC/implementations/controller_classes/controller_MPI.py``: ``run`` :90-168, ``restart_block`` :170-216,
``send_full / recv_full`` :235-305, stages ``spread`` :454, ``predict`` :481, ``it_check`` :552, ``it_fine`` :635,
``it_down`` :671, ``it_coarse`` :713, ``it_up`` :754) and the ring protocol of its convergence check
(``convergence_controller_classes/check_convergence.py:114-160``), with the mpi4py calls replaced by NCCL send/recv of
device fields over NVLink (``parallel.TorchComm``):

* step-to-step hand-over ``uend -> u[0]`` of the next slice on every level: ``mesh.isend`` / ``mesh.irecv``
  (controller_MPI.py:228, 267) followed by ``eval_f(u[0])`` on the receiver (:233);
* convergence status: one scalar along the ring, or an all-reduce when ``all_to_done`` is set;
* block hand-over: broadcast of the last slice's ``uend`` (:125-131).

Not implemented (the reference's optional machinery around the path): the iteration estimator with interruptible
waits, adaptive step sizes / restarts, per-step varying ``dt``.  A communicator of size 1 (``parallel.LocalComm``) gives
plain serial SDC / MLSDC.
"""
import logging

import numpy as np

from .comm import LAND, LOR
from .controller import DefaultHooks
from .core import Bag, Step
from .errors import ControllerError
from .parallel import LocalComm


class controller_MPI:
    def __init__(self, controller_params, description, comm=None):
        self.params = Bag(logger_level=20, hook_class=[], all_to_done=False, predict_type=None, mssdc_jac=True,
                          dump_setup=False, fname="run_pid.log", use_iteration_estimator=False)
        for k, v in controller_params.items():
            setattr(self.params, k, v)
        if self.params.use_iteration_estimator:
            raise ControllerError("the iteration estimator is not implemented in the device controller")
        self.logger = logging.getLogger("controller")
        self.logger.setLevel(self.params.logger_level)
        self.comm = comm if comm is not None else LocalComm()
        self.S = Step(description)
        self.S.status.time_size = self.comm.Get_size()
        hook_classes = self.params.hook_class if isinstance(self.params.hook_class, list) else [self.params.hook_class]
        self.hooks = [DefaultHooks()] + [h() for h in hook_classes]
        nlev = len(self.S.levels)
        if self.comm.Get_size() > 1 and nlev > 1:
            for L in self.S.levels:
                if not L.sweep.coll.right_is_node or L.sweep.params.do_coll_update:
                    raise ControllerError("For PFASST to work, we assume uend^k = u_M^k")
        if nlev == 1 and self.params.predict_type is not None:
            self.logger.warning("you have specified a predictor type but only a single level.. predictor will be ignored")
        self.req_send = [None] * nlev

    # ---- plumbing -----------------------------------------------------------------------------------------------------
    def _call(self, name, level=0, **kw):
        for h in self.hooks:
            getattr(h, name)(step=self.S, level_number=level, **kw)

    def return_stats(self):
        stats = {}
        for h in self.hooks:
            stats.update(h.return_stats())
        return stats

    # ---- driver (controller_MPI.py:90-168) --------------------------------------------------------------------------
    def run(self, u0, t0, Tend):
        S, comm = self.S, self.comm
        for h in self.hooks:
            h.reset_stats()
        eps = 10 * np.finfo(float).eps
        dt = S.dt
        all_dt = comm.allgather(dt)
        if any(abs(d - dt) > eps for d in all_dt):
            raise ControllerError("the device controller needs the same dt on every time slice")
        size, rank = comm.Get_size(), comm.Get_rank()
        # block structure is known up front (fixed dt): full blocks of `size` slices and possibly a shorter last one.
        # Sub-communicators are created collectively NOW, while every rank is still here.
        nsteps = 0
        while t0 + nsteps * dt < Tend - eps:
            nsteps += 1
        if nsteps == 0:
            raise ControllerError("Nothing to do, check t0, dt and Tend!")
        last = nsteps % size
        comm_last = comm.first(last) if last else None
        nblocks = (nsteps + size - 1) // size

        uend, tend = u0, t0
        S.status.slot = rank
        self._call("post_setup")
        self._call("pre_run")
        for blk in range(nblocks):
            nact = size if (blk < nblocks - 1 or last == 0) else last
            if rank >= nact:
                break
            c = comm if nact == size else comm_last
            S.status.slot = c.Get_rank()
            self.restart_block(nact, tend + S.status.slot * dt, uend)
            while not S.status.done:
                self.pfasst(c, nact)
            # hand the end value of the last slice to everybody as the next initial value (:125-131)
            S.levels[0].uend.bcast(root=nact - 1, comm=c)
            uend = S.levels[0].uend
            tend = c.bcast(S.time + S.dt, root=nact - 1)
        self._call("post_run")
        return uend, self.return_stats()

    def restart_block(self, size, time, u0):
        """controller_MPI.py:170-216."""
        S = self.S
        S.prev = (S.status.slot - 1) % size
        S.next = (S.status.slot + 1) % size
        S.reset_step()
        S.status.first = S.prev == size - 1
        S.status.last = S.next == 0
        S.init_step(u0)
        S.status.done = False
        S.status.iter = 0
        S.status.stage = "SPREAD"
        S.status.prev_done = False
        S.status.force_done = False
        S.status.force_continue = False
        S.status.time_size = size
        self.req_send = [None] * len(S.levels)
        for L in S.levels:
            L.tag = None
            L.status.time = time
            L.status.sweep = 1

    # ---- communication (controller_MPI.py:218-305) -----------------------------------------------------------------
    def recv(self, target, source, tag, comm):
        target.u[0].irecv(source=source, tag=tag, comm=comm).Wait()
        target.f[0] = target.prob.eval_f(target.u[0], target.time)

    def send_full(self, comm, blocking=False, level=0, add_to_stats=False):
        S = self.S
        self._call("pre_comm", level)
        if not blocking and self.req_send[level] is not None:
            self.req_send[level].Wait()
        S.levels[level].sweep.compute_end_point()
        if not S.status.last:
            self.req_send[level] = S.levels[level].uend.isend(dest=S.next, tag=level * 100 + S.status.iter, comm=comm)
            if blocking:
                self.req_send[level].Wait()
        self._call("post_comm", level, add_to_stats=add_to_stats)

    def recv_full(self, comm, level=0, add_to_stats=False):
        S = self.S
        self._call("pre_comm", level)
        if not S.status.first and not S.status.prev_done:
            self.recv(S.levels[level], S.prev, level * 100 + S.status.iter, comm)
        self._call("post_comm", level, add_to_stats=add_to_stats)

    # ---- convergence (check_convergence.py:60-160) -------------------------------------------------------------------
    def check_convergence(self, comm):
        S = self.S
        L = S.levels[0]
        iter_converged = S.status.iter >= S.params.maxiter
        res_converged = L.status.residual <= L.params.restol and (S.status.iter > 0 or L.status.sweep > 0)
        S.status.done = bool((iter_converged or res_converged or S.status.force_done) and not S.status.force_continue)
        if self.params.all_to_done:
            self._call("pre_comm")
            S.status.done = comm.allreduce(S.status.done, op=LAND)
            S.status.force_done = comm.allreduce(S.status.force_done, op=LOR)
            self._call("post_comm", add_to_stats=True)
            S.status.done = S.status.done or S.status.force_done
        else:
            self._call("pre_comm")
            if not S.status.first and not S.status.prev_done:
                S.status.prev_done = bool(comm.recv_scalar(source=S.status.slot - 1))
                S.status.done = S.status.done and S.status.prev_done
            if not S.status.last:
                comm.send_scalar(S.status.done, dest=S.status.slot + 1)
            self._call("post_comm", add_to_stats=True)
        S.status.force_continue = False

    # ---- stages --------------------------------------------------------------------------------------------------------
    def pfasst(self, comm, num_procs):
        stage = self.S.status.stage
        fn = {"SPREAD": self.spread, "PREDICT": self.predict, "IT_CHECK": self.it_check, "IT_FINE": self.it_fine,
              "IT_DOWN": self.it_down, "IT_COARSE": self.it_coarse, "IT_UP": self.it_up}.get(stage)
        if fn is None:
            raise ControllerError("Weird stage, got %s" % stage)
        fn(comm, num_procs)

    def spread(self, comm, num_procs):
        S = self.S
        self._call("pre_step")
        S.levels[0].sweep.predict()
        S.status.stage = "PREDICT" if len(S.levels) > 1 else "IT_CHECK"

    def predict(self, comm, num_procs):
        S = self.S
        nlev = len(S.levels)
        self._call("pre_predict")
        ptype = self.params.predict_type
        if ptype is None:
            pass
        elif ptype == "fine_only":
            S.levels[0].sweep.update_nodes()
        elif ptype == "pfasst_burnin":
            for l in range(1, nlev):
                S.transfer(source=S.levels[l - 1], target=S.levels[l])
            for p in range(S.status.slot + 1):
                if p != 0:
                    self.recv_full(comm, level=nlev - 1)
                S.levels[-1].sweep.update_nodes()
                S.levels[-1].sweep.compute_end_point()
                self.send_full(comm, blocking=True, level=nlev - 1, add_to_stats=(p == S.status.slot))
            for l in range(nlev - 1, 0, -1):
                S.transfer(source=S.levels[l], target=S.levels[l - 1])
            self.send_full(comm, level=0)
            self.recv_full(comm, level=0)
            S.levels[0].sweep.update_nodes()
        elif ptype == "fmg":
            raise NotImplementedError("FMG predictor is not yet implemented")
        else:
            raise ControllerError("Wrong predictor type, got %s" % ptype)
        self._call("post_predict")
        S.status.stage = "IT_CHECK"

    def it_check(self, comm, num_procs):
        S = self.S
        self.send_full(comm, level=0)
        self.recv_full(comm, level=0)
        S.levels[0].sweep.compute_residual(stage="IT_CHECK")
        if S.status.iter > 0:
            self._call("post_iteration")
        self.check_convergence(comm)
        if not S.status.done:
            S.status.iter += 1
            self._call("pre_iteration")
            if len(S.levels) > 1:
                S.status.stage = "IT_DOWN"
            elif num_procs == 1 or self.params.mssdc_jac:
                S.status.stage = "IT_FINE"
            else:
                S.status.stage = "IT_COARSE"
        else:
            for req in self.req_send:
                if req is not None:
                    req.Wait()
            self._call("post_step")
            S.status.stage = "DONE"

    def _sweep_level(self, comm, l, stage, k=None, add_to_stats=False):
        S = self.S
        self.send_full(comm, level=l)
        self.recv_full(comm, level=l, add_to_stats=add_to_stats)
        self._call("pre_sweep", l)
        if k is not None:
            S.levels[l].sweep.updateVariableCoeffs(k + 1)
        S.levels[l].sweep.update_nodes()
        S.levels[l].sweep.compute_residual(stage=stage)
        self._call("post_sweep", l)

    def it_fine(self, comm, num_procs):
        S = self.S
        nsweeps = S.levels[0].params.nsweeps
        S.levels[0].status.sweep = 0
        for k in range(nsweeps):
            S.levels[0].status.sweep += 1
            self._sweep_level(comm, 0, "IT_FINE", k=k, add_to_stats=(k == nsweeps - 1))
        S.status.stage = "IT_CHECK"

    def it_down(self, comm, num_procs):
        S = self.S
        S.transfer(source=S.levels[0], target=S.levels[1])
        for l in range(1, len(S.levels) - 1):
            for _ in range(S.levels[l].params.nsweeps):
                self._sweep_level(comm, l, "IT_DOWN")
            S.transfer(source=S.levels[l], target=S.levels[l + 1])
        S.status.stage = "IT_COARSE"

    def it_coarse(self, comm, num_procs):
        S = self.S
        lc = len(S.levels) - 1
        self.recv_full(comm, level=lc)
        self._call("pre_sweep", lc)
        if S.levels[-1].params.nsweeps != 1:
            raise ControllerError("this controller can only work with one sweep on the coarse level")
        S.levels[-1].sweep.update_nodes()
        S.levels[-1].sweep.compute_residual(stage="IT_COARSE")
        self._call("post_sweep", lc)
        S.levels[-1].sweep.compute_end_point()
        self.send_full(comm, blocking=True, level=lc, add_to_stats=True)
        S.status.stage = "IT_UP" if len(S.levels) > 1 else "IT_CHECK"

    def it_up(self, comm, num_procs):
        S = self.S
        for l in range(len(S.levels) - 1, 0, -1):
            S.transfer(source=S.levels[l], target=S.levels[l - 1])
            if l - 1 > 0:
                nsweeps = S.levels[l - 1].params.nsweeps
                for k in range(nsweeps):
                    self._sweep_level(comm, l - 1, "IT_UP", add_to_stats=(k == nsweeps - 1))
        S.status.stage = "IT_FINE"
