"""Exception types.  When pySDC itself is importable its exception classes are re-exported, so that code written
against ``pySDC.core.errors`` (core/errors.py) catches what the B200 classes raise; otherwise stand-alone classes
with the same names are defined."""
try:  # pragma: no cover - depends on the environment
    from pySDC.core.errors import (  # noqa: F401
        CollocationError, ControllerError, ParameterError, ProblemError, TransferError, UnlockError,
    )
except Exception:  # pySDC (or its qmat dependency) not installed

    class ParameterError(Exception):
        """A parameter is missing or has an unusable value."""

    class ProblemError(Exception):
        """A problem class was set up inconsistently."""

    class CollocationError(Exception):
        """The collocation set-up is invalid."""

    class UnlockError(Exception):
        """Data of a level was used before a predictor unlocked it."""

    class ControllerError(Exception):
        """The controller reached an inconsistent state."""

    class TransferError(Exception):
        """A transfer class cannot work with the given levels / data."""


class BackendError(RuntimeError):
    """libsdcb200.so is missing, failed to load, or a kernel call returned an error.  There is no CPU fallback."""
