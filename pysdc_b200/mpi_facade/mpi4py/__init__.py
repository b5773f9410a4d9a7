"""A minimal ``mpi4py`` look-alike on ``torch.distributed`` (NCCL for device fields, object collectives for Python
values), so that the reference's ``controller_MPI`` and MPI-aware convergence controllers run UNMODIFIED on a machine
without MPI:

    import sys, pysdc_b200.mpi_facade; sys.path.insert(0, pysdc_b200.mpi_facade.PATH)   # before importing pySDC
    from mpi4py import MPI                                                            # resolves to this package
    from pySDC.implementations.controller_classes.controller_MPI import controller_MPI
    controller_MPI(controller_params, description, comm=MPI.COMM_WORLD)

Only what pySDC's time-parallel control path uses is provided (controller_classes/controller_MPI.py,
core/convergence_controller.py:360-452, convergence_controller_classes/{check_convergence, basic_restarting,
spread_step_sizes}.py, datatype ``isend/irecv/bcast``): see ``MPI.Intracomm``.
"""
from . import MPI  # noqa: F401
