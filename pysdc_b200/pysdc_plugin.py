"""Drop-in classes for an unmodified pySDC installation.

    from pysdc_b200.pysdc_plugin import heatNd_unforced, generic_implicit     # instead of pySDC.implementations...
    description = {'problem_class': heatNd_unforced, 'sweeper_class': generic_implicit, ...}   # everything else as is
    controller_nonMPI(num_procs=1, controller_params=..., description=description).run(u0, t0, Tend)

The numerical mix-ins of ``sweepers.py`` / ``problems.py`` are bound to pySDC's OWN base classes
(``pySDC.core.sweeper.Sweeper``, ``pySDC.core.problem.Problem``), so collocation and QDelta coefficients come from
pySDC / qmat, ``Level`` and ``Step`` type checks pass (core/sweeper.py:253-256), and pySDC's controllers, hooks and
convergence controllers run around them unchanged.  Importing this module needs pySDC (and its qmat dependency).
"""
from pySDC.core.problem import Problem as _PySDCProblem
from pySDC.core.sweeper import Sweeper as _PySDCSweeper

from . import problems as _problems
from . import sweepers as _sweepers
from .datatypes import comp2_mesh, imex_mesh, mesh  # noqa: F401
from .transfer import mesh_to_mesh  # noqa: F401  (space_transfer_class for multi-level runs)

# the reference's CuPy datatypes (datatype_classes/cupy_mesh.py), for scripts written against its GPU problem classes
# (problem_classes/HeatEquation_ND_FD_CuPy.py, AllenCahn_2D_FD_gpu.py: same class names as the CPU ones)
cupy_mesh, imex_cupy_mesh, comp2_cupy_mesh = mesh, imex_mesh, comp2_mesh

globals().update({k: v for k, v in _problems._bind(_PySDCProblem).items()})
globals().update({k: v for k, v in _sweepers._bind(_PySDCSweeper).items()})

__all__ = ["mesh", "imex_mesh", "comp2_mesh", "cupy_mesh", "imex_cupy_mesh", "comp2_cupy_mesh", "heatNd_unforced", "heatNd_forced", "allencahn_fullyimplicit", "allencahn_semiimplicit",
           "allencahn_semiimplicit_v2", "allencahn_multiimplicit", "allencahn_multiimplicit_v2", "multi_implicit", "generic_implicit", "imex_1st_order", "generic_implicit_MPI", "imex_1st_order_MPI",
           "mesh_to_mesh"]
