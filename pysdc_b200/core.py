"""Stand-alone host-side core: the small part of pySDC's object model the sweep path touches.

The hot-path classes of this package (``sweepers.py``, ``problems.py``) are written as mix-ins over an abstract base;
this module provides that base *without* pySDC, with the same attribute and method names as the reference
(``pySDC/core/{common,problem,sweeper,level,step}.py``), so that the package runs on a machine that has no pySDC
(the GPU box) and the tests read like the reference's.  ``pysdc_plugin.py`` binds the same mix-ins to pySDC's own
base classes for use inside an unmodified pySDC installation.
"""
import logging

import numpy as np

from .errors import ParameterError
from .quadrature import CollBase, make_qdelta_generator


class Bag:
    """Attribute bag that refuses unknown names once frozen (role of helpers/pysdc_helper.py FrozenClass)."""

    def __init__(self, **defaults):
        object.__setattr__(self, "_frozen", False)
        for k, v in defaults.items():
            setattr(self, k, v)

    def _freeze(self):
        object.__setattr__(self, "_frozen", True)

    def __setattr__(self, key, value):
        if self._frozen and key not in self.__dict__:
            raise TypeError(f"{type(self).__name__!r} is frozen, cannot add attribute {key!r}")
        object.__setattr__(self, key, value)

    def get(self, key, default=None):
        return self.__dict__.get(key, default)


class WorkCounter:
    """core/problem.py:16-40."""

    def __init__(self):
        self.niter = 0

    def __call__(self, *args, **kwargs):
        self.niter += 1

    def decrement(self):
        self.niter -= 1

    def __str__(self):
        return f"{self.niter}"


class ReadOnlyError(Exception):
    pass


class Problem:
    """core/problem.py:43-215 and core/common.py:25-79 (parameter registration)."""

    logger = logging.getLogger("problem")
    dtype_u = None
    dtype_f = None

    def __init__(self, init):
        object.__setattr__(self, "_par_names", set())
        object.__setattr__(self, "_par_names_ro", set())
        self.work_counters = {}
        self.init = init

    def _makeAttributeAndRegister(self, *names, localVars=None, readOnly=False):
        if names and localVars is None:
            raise ValueError("a dictionary must be provided in localVars with parameters values")
        for name in names:
            if name not in localVars:
                raise ValueError(f"value for {name} not given in localVars")
            object.__setattr__(self, name, localVars[name])
        (self._par_names_ro if readOnly else self._par_names).update(names)

    @property
    def params(self):
        return {name: getattr(self, name) for name in self._par_names | self._par_names_ro}

    def __setattr__(self, name, value):
        if name in self.__dict__.get("_par_names_ro", ()):
            raise ReadOnlyError(name)
        object.__setattr__(self, name, value)

    @property
    def u_init(self):
        return self.dtype_u(self.init)

    @property
    def f_init(self):
        return self.dtype_f(self.init)

    @classmethod
    def get_default_sweeper_class(cls):
        raise NotImplementedError(f"No default sweeper class implemented for {cls} problem!")

    def eval_f(self, u, t):
        raise NotImplementedError

    def solve_system(self, rhs, factor, u0, t):
        raise NotImplementedError


class Sweeper:
    """core/sweeper.py:55-276: collocation + QDelta set-up; the numerical methods live in ``sweepers.py``."""

    def __init__(self, params, level):
        self.logger = logging.getLogger("sweeper")
        if "num_nodes" not in params:
            raise ParameterError("need num_nodes to instantiate step, only got %s" % str(params.keys()))
        coll_class = params.get("collocation_class", CollBase)
        if params.get("initial_guess", "spread") == "random":  # core/sweeper.py:77-79
            params["random_seed"] = params.get("random_seed", 1984)
            self.rng = np.random.RandomState(params["random_seed"])
        self.params = Bag(do_coll_update=False, initial_guess="spread", skip_residual_computation=())
        for k, v in params.items():
            if k != "collocation_class":
                setattr(self.params, k, v)
        self.params._freeze()
        self.coll = coll_class(**{k: v for k, v in params.items() if k != "collocation_class"})
        if not self.coll.right_is_node and not self.params.do_coll_update:
            self.logger.warning("we need to do a collocation update here, since the right end point is not a node. "
                                "Changing this!")
            self.params.do_coll_update = True
        self.__level = level
        self.parallelizable = False
        self.genQI = self.genQE = None

    def _check_lower(self, QD, strict):
        if np.any(np.triu(QD, k=0 if strict else 1) != 0.0):
            raise ParameterError("Strictly lower triangular matrix expected!" if strict
                                 else "Lower triangular matrix expected!")
        if np.allclose(np.diag(np.diag(QD)), QD):
            self.parallelizable = True

    def get_Qdelta_implicit(self, qd_type, k=None):
        if self.genQI is None or self.genQI_name != qd_type:
            self.genQI, self.genQI_name = make_qdelta_generator(qd_type, self.coll), qd_type
        QD = np.zeros_like(self.coll.Qmat)
        QD[1:, 1:] = self.genQI.coeffs(k)
        self._check_lower(QD, strict=False)
        return QD

    def get_Qdelta_explicit(self, qd_type, k=None):
        if self.genQE is None or self.genQE_name != qd_type:
            self.genQE, self.genQE_name = make_qdelta_generator(qd_type, self.coll), qd_type
        QD = np.zeros_like(self.coll.Qmat)
        QD[1:, 1:] = self.genQE.coeffs(k)
        QD[1:, 0] = self.genQE.dtau(k)
        self._check_lower(QD, strict=True)
        return QD

    def updateVariableCoeffs(self, k):
        if self.genQI is not None and self.genQI.isKDependent():
            self.QI = self.get_Qdelta_implicit(self.genQI_name, k=k)
        if self.genQE is not None and self.genQE.isKDependent():
            self.QE = self.get_Qdelta_explicit(self.genQE_name, k=k)

    @property
    def level(self):
        return self.__level

    @level.setter
    def level(self, L):
        assert isinstance(L, Level)
        self.__level = L

    @property
    def rank(self):
        return 0


class Level:
    """core/level.py:42-191."""

    def __init__(self, problem_class, problem_params, sweeper_class, sweeper_params, level_params, level_index):
        self.params = Bag(dt=None, dt_initial=None, restol=-1.0, nsweeps=1, residual_type="full_abs")
        for k, v in level_params.items():
            setattr(self.params, k, v)
        self.params._freeze()
        self.params.dt_initial = self.params.dt * 1.0 if self.params.dt is not None else None
        self.status = self._new_status()
        self.__sweep = sweeper_class(sweeper_params, self)
        self.__prob = problem_class(**problem_params)
        self.level_index = level_index
        M = self.sweep.coll.num_nodes
        self.uend = None
        self.u = [None] * (M + 1)
        self.uold = [None] * (M + 1)
        self.f = [None] * (M + 1)
        self.fold = [None] * (M + 1)
        self.tau = [None] * M
        self.residual = [None] * M
        self.tag = None

    @staticmethod
    def _new_status():
        s = Bag(residual=None, unlocked=False, updated=False, time=None, dt_new=None, sweep=None)
        s._freeze()
        return s

    def reset_level(self, reset_status=True):
        if reset_status:
            self.status = self._new_status()
        M = self.sweep.coll.num_nodes
        self.uend = None
        self.u = [None] * (M + 1)
        self.uold = [None] * (M + 1)
        self.f = [None] * (M + 1)
        self.fold = [None] * (M + 1)
        self.tau = [None] * M

    @property
    def sweep(self):
        return self.__sweep

    @property
    def prob(self):
        return self.__prob

    @property
    def time(self):
        return self.status.time

    @property
    def dt(self):
        return self.params.dt


class Step:
    """core/step.py:87-331 (level hierarchy, per-level parameter lists, init/reset)."""

    def __init__(self, description):
        for key in ("problem_class", "sweeper_class", "sweeper_params", "level_params"):
            if key not in description:
                raise ParameterError(f"need {key} to instantiate step, only got {list(description)}")
        self.params = Bag(maxiter=None, errtol=None)
        for k, v in description.get("step_params", {}).items():
            setattr(self.params, k, v)
        self.params._freeze()
        self.status = Bag(iter=None, stage=None, slot=None, first=None, last=None, pred_cnt=None, done=None,
                          force_done=None, force_continue=False, prev_done=None, time_size=None, diff_old_loc=None,
                          diff_first_loc=None, restart=False)
        self.levels = []
        self.base_transfer = None
        self._transfer = {}
        self.prev = self.next = None
        plist = self._to_list(description.get("problem_params", {}))
        slist = self._to_list(description["sweeper_params"])
        llist = self._to_list(description["level_params"])
        nlev = max(len(plist), len(slist), len(llist))
        expand = lambda lst: lst if len(lst) == nlev else lst * nlev if len(lst) == 1 else None  # noqa: E731
        plist, slist, llist = expand(plist), expand(slist), expand(llist)
        if None in (plist, slist, llist):
            raise ParameterError("per-level parameter lists have inconsistent lengths")
        for l in range(nlev):
            self.levels.append(Level(description["problem_class"], dict(plist[l]), description["sweeper_class"],
                                     dict(slist[l]), dict(llist[l]), l))
        if nlev > 1:
            if "space_transfer_class" not in description:
                raise ParameterError("need space_transfer_class to instantiate step with more than one level")
            from .transfer import BaseTransfer
            for l in range(nlev - 1):
                bt = BaseTransfer(self.levels[l], self.levels[l + 1], description.get("base_transfer_params", {}),
                                  description["space_transfer_class"], description.get("space_transfer_params", {}))
                self.base_transfer = bt
                self._transfer[(l, l + 1)] = bt.restrict
                self._transfer[(l + 1, l)] = bt.prolong_f if bt.params.finter else bt.prolong

    @staticmethod
    def _to_list(d):
        """Dict whose values may be per-level lists -> list of per-level dicts (core/step.py:174-199)."""
        lens = {len(v) for v in d.values() if isinstance(v, list)}
        if len(lens) > 1:
            raise ParameterError("per-level parameter lists have inconsistent lengths")
        n = lens.pop() if lens else 1
        return [{k: (v[i] if isinstance(v, list) else v) for k, v in d.items()} for i in range(n)]

    def transfer(self, source, target):
        self._transfer[(source.level_index, target.level_index)]()

    def reset_step(self):
        for L in self.levels:
            L.reset_level()

    def init_step(self, u0):
        P = self.levels[0].prob
        self.levels[0].u[0] = P.dtype_u(u0)

    @property
    def dt(self):
        return self.levels[0].dt

    @property
    def time(self):
        return self.levels[0].time
