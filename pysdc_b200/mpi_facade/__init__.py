"""Directory to put on ``sys.path`` so that ``import mpi4py`` resolves to the torch.distributed-backed facade."""
import os

PATH = os.path.dirname(os.path.abspath(__file__))
