"""pysdc_b200 — B200-native (sm_100a) implementation of pySDC's SDC sweep hot path.

Public surface (same names as the reference classes they replace):

* datatypes  ``mesh``, ``imex_mesh``                                  (``pysdc_b200.datatypes``)
* problems   ``heatNd_unforced``, ``heatNd_forced``, ``allencahn_fullyimplicit``   (``pysdc_b200.problems``)
* sweepers   ``generic_implicit``, ``imex_1st_order``                  (``pysdc_b200.sweepers``)
* stand-alone controller ``controller_nonMPI`` + ``Step`` / ``Level``  (``pysdc_b200.controller``, ``pysdc_b200.core``)
* ``pysdc_b200.pysdc_plugin`` — the same classes bound to pySDC's own base classes (needs pySDC importable)

All numerical work runs in ``pysdc_b200/lib/libsdcb200.so`` (CUDA, C ABI in ``include/sdc_b200.h``); importing this
package does not need a GPU, using it does.
"""
__version__ = "0.1.0"

from .errors import BackendError, ParameterError, ProblemError  # noqa: F401
