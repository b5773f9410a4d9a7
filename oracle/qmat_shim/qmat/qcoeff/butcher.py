"""Only present so that ``pySDC/implementations/sweeper_classes/Runge_Kutta.py:3`` can be imported;
the Runge-Kutta sweepers are outside the SDC sweep path and are not restated."""


class _Unavailable(dict):
    def __missing__(self, key):
        raise NotImplementedError(f"RK scheme {key!r}: Butcher tables are not part of the qmat stand-in")


RK_SCHEMES = _Unavailable()
