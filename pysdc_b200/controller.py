"""Host controller for the stand-alone mode: serial single-level SDC time stepping with the stage sequence of the
reference's ``controller_nonMPI`` (``pySDC/implementations/controller_classes/controller_nonMPI.py:85-178, 297-580``:
SPREAD -> IT_CHECK <-> IT_FINE -> DONE), its convergence test (``convergence_controller_classes/
check_convergence.py:60-112``), hook call points (``core/hooks.py:106-245``) and statistics layout.

Inside an unmodified pySDC installation this module is not needed: pySDC's own controllers drive the classes exported
by ``pysdc_plugin``.  Multi-level / time-parallel runs (MLSDC, PFASST) are in ``pfasst.py``.
"""
import logging
import time as _time

import numpy as np

from .core import Bag, Step
from .errors import ControllerError
from .stats import Entry


class Hooks:
    """Hook base class: same call points and ``add_to_stats`` signature as ``pySDC/core/hooks.py``."""

    def __init__(self):
        self.stats = {}
        self.logger = logging.getLogger("hooks")

    def add_to_stats(self, value, process=-1, process_sweeper=-1, time=-1, level=-1, iter=-1, sweep=-1, type=-1,
                     num_restarts=-1, **kwargs):
        self.stats[Entry(process, process_sweeper, time, level, iter, sweep, type, num_restarts)] = value

    def reset_stats(self):
        self.stats = {}

    def return_stats(self):
        return self.stats

    def pre_setup(self, step, level_number): pass
    def post_setup(self, step, level_number): pass
    def pre_run(self, step, level_number): pass
    def post_run(self, step, level_number): pass
    def pre_step(self, step, level_number): pass
    def post_step(self, step, level_number): pass
    def pre_predict(self, step, level_number): pass
    def post_predict(self, step, level_number): pass
    def pre_iteration(self, step, level_number): pass
    def post_iteration(self, step, level_number): pass
    def pre_sweep(self, step, level_number): pass
    def post_sweep(self, step, level_number): pass
    def pre_comm(self, step, level_number): pass
    def post_comm(self, step, level_number, add_to_stats=False): pass


class DefaultHooks(Hooks):
    """Residual / iteration-count logging of ``implementations/hooks/default_hook.py`` plus wall-clock timings of
    ``hooks/log_timings.py`` (``timing_run``, ``timing_step``)."""

    def _common(self, step, L):
        return dict(process=step.status.slot, process_sweeper=L.sweep.rank, time=L.time, level=L.level_index,
                    iter=step.status.iter, sweep=L.status.sweep)

    def pre_run(self, step, level_number):
        self._t_run = _time.perf_counter()

    def post_run(self, step, level_number):
        self.add_to_stats(_time.perf_counter() - self._t_run, process=step.status.slot, time=-1, level=-1, iter=-1,
                          sweep=-1, type="timing_run")

    def pre_step(self, step, level_number):
        self._t_step = _time.perf_counter()

    def post_sweep(self, step, level_number):
        L = step.levels[level_number]
        self.add_to_stats(L.status.residual, type="residual_post_sweep", **self._common(step, L))

    def post_iteration(self, step, level_number):
        L = step.levels[level_number]
        self.add_to_stats(L.status.residual, type="residual_post_iteration", **self._common(step, L))

    def post_step(self, step, level_number):
        L = step.levels[level_number]
        c = self._common(step, L)
        self.add_to_stats(_time.perf_counter() - self._t_step, type="timing_step", **c)
        self.add_to_stats(step.status.iter, type="niter", **c)
        self.add_to_stats(L.status.residual, type="residual_post_step", **c)


class LogWork(Hooks):
    """Per-step increments of the problem's work counters (``implementations/hooks/log_work.py:4-55``)."""

    def pre_step(self, step, level_number):
        self._before = [{k: c.niter for k, c in L.prob.work_counters.items()} for L in step.levels]

    def post_step(self, step, level_number):
        L = step.levels[level_number]
        for key, before in self._before[level_number].items():
            self.add_to_stats(L.prob.work_counters[key].niter - before, process=step.status.slot,
                              process_sweeper=L.sweep.rank, time=L.time + L.dt, level=L.level_index,
                              iter=step.status.iter, sweep=L.status.sweep, type=f"work_{key}")


class LogToFile(Hooks):
    """Solutions to a FieldsIO file every ``time_increment`` time units, as ``implementations/hooks/log_solution.py:
    207-282`` does: the problem provides the file (``getOutputFile``) and the host copy (``processSolutionForOutput``,
    an asynchronous device -> pinned-host copy here); an existing file is re-opened and continued when the run starts
    at t > 0."""

    filename = "myRun.pySDC"
    time_increment = 0
    counter = 0

    def __init__(self):
        super().__init__()
        self.outfile, self.t_next_log = None, 0

    def pre_run(self, step, level_number):
        import os

        L = step.levels[level_number]
        if os.path.isfile(self.filename) and L.time > 0:
            self.outfile = L.prob.openOutputFile(self.filename)
        else:
            self.outfile = L.prob.getOutputFile(self.filename)
            self.outfile.addField(time=L.time, field=L.prob.processSolutionForOutput(L.u[0]))
        type(self).counter = len(self.outfile.times)

    def post_step(self, step, level_number):
        L = step.levels[level_number]
        if self.t_next_log == 0:
            self.t_next_log = L.time + self.time_increment
        if L.time + L.dt >= self.t_next_log:
            self.outfile.addField(time=L.time + L.dt, field=L.prob.processSolutionForOutput(L.uend))
            self.t_next_log = max([L.time + L.dt, self.t_next_log]) + self.time_increment
            type(self).counter += 1  # (not len(outfile.times): that would wait for the copy that is still in flight)

    def post_run(self, step, level_number):
        self.outfile.flush()

    @classmethod
    def load(cls, index):
        from .fields_io import RectilinearFile

        t, u = RectilinearFile.fromFile(cls.filename).readField(index)
        return {"t": t, "u": u}


class controller_nonMPI:
    """``controller_nonMPI(num_procs, controller_params, description).run(u0, t0, Tend) -> (uend, stats)``."""

    def __init__(self, num_procs, controller_params, description):
        self.params = Bag(logger_level=20, hook_class=[], all_to_done=False, predict_type=None, mssdc_jac=True,
                          dump_setup=False, fname="run_pid.log", use_iteration_estimator=False)
        for k, v in controller_params.items():
            setattr(self.params, k, v)
        self.logger = logging.getLogger("controller")
        self.logger.setLevel(self.params.logger_level)
        if num_procs != 1:
            raise ControllerError("time-parallel runs use pysdc_b200.pfasst; this controller steps serially")
        self.MS = [Step(description)]
        if len(self.MS[0].levels) > 1:
            raise ControllerError("multi-level runs use pysdc_b200.pfasst")
        hook_classes = self.params.hook_class if isinstance(self.params.hook_class, list) else [self.params.hook_class]
        self.hooks = [DefaultHooks()] + [h() for h in hook_classes]
        self.nsweeps = [L.params.nsweeps for L in self.MS[0].levels]

    def _call(self, name, S, level=0, **kw):
        for h in self.hooks:
            getattr(h, name)(step=S, level_number=level, **kw)

    def return_stats(self):
        stats = {}
        for h in self.hooks:
            stats.update(h.return_stats())
        return stats

    @staticmethod
    def check_convergence(S):
        """check_convergence.py:60-92."""
        L = S.levels[0]
        iter_converged = S.status.iter >= S.params.maxiter
        res_converged = L.status.residual <= L.params.restol and (S.status.iter > 0 or L.status.sweep > 0)
        return bool((iter_converged or res_converged or S.status.force_done) and not S.status.force_continue)

    def run(self, u0, t0, Tend):
        S = self.MS[0]
        L = S.levels[0]
        for h in self.hooks:
            h.reset_stats()
        t = t0
        if not t < Tend - 10 * np.finfo(float).eps:
            raise ControllerError("Nothing to do, check t0, dt and Tend.")
        S.status.slot = 0
        # the first block is set up before the pre-run hooks fire (controller_nonMPI.py:101-110): they may read L.u[0]
        S.reset_step()
        S.init_step(u0)
        L.status.time = t0
        self._call("post_setup", S)
        self._call("pre_run", S)
        uend = u0
        while t < Tend - 10 * np.finfo(float).eps:  # controller_nonMPI.py:112,164
            # restart_block (:180-224)
            S.reset_step()
            S.status.first = S.status.last = True
            S.init_step(uend)
            S.status.done = False
            S.status.iter = 0
            S.status.force_done = False
            S.status.stage = "SPREAD"
            L.status.sweep = 1
            L.status.time = t
            # SPREAD (:334-357)
            self._call("pre_step", S)
            L.sweep.predict()
            S.status.stage = "IT_CHECK"
            while True:
                # IT_CHECK (:479-543)
                L.sweep.compute_residual(stage="IT_CHECK")
                if S.status.iter > 0:
                    self._call("post_iteration", S)
                S.status.done = self.check_convergence(S)
                S.status.force_continue = False
                if S.status.done:
                    L.sweep.compute_end_point()
                    self._call("post_step", S)
                    S.status.stage = "DONE"
                    break
                S.status.iter += 1
                self._call("pre_iteration", S)
                # IT_FINE (:545-580)
                L.status.sweep = 0
                for k in range(self.nsweeps[0]):
                    L.status.sweep += 1
                    self._call("pre_sweep", S)
                    L.sweep.updateVariableCoeffs(k + 1)
                    L.sweep.update_nodes()
                    L.sweep.compute_residual(stage="IT_FINE")
                    self._call("post_sweep", S)
            uend = L.uend
            t = t + L.dt
        self._call("post_run", S)
        return uend, self.return_stats()
