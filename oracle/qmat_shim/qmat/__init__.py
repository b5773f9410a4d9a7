"""Stand-in for the third-party ``qmat`` package (PyPI ``qmat>=0.1.19``, pinned by the
reference in ``pyproject.toml:34``; conda envs pin ``>=0.1.8``, ``etc/environment-base.yml:16``).

TEST INFRASTRUCTURE ONLY.  The reference imports ``qmat`` at module scope
(``pySDC/core/collocation.py:4``, ``pySDC/core/sweeper.py:4``, ``pySDC/core/base_transfer.py:9``)
and the package is neither installed nor vendored in this image, so the reference cannot be
imported without it.  This module restates the *published* algorithms of qmat for exactly the
call sites the SDC sweep path touches; nothing here is shipped in the product path
(``pysdc_b200`` has its own, independently written quadrature module, and a CPU test checks
that the two agree to round-off).

Parity status: the reference's own tests pin this boundary by *properties* only
(``pySDC/tests/test_collocation.py:19-120``, ``tests/test_sweepers/test_preconditioners.py:15-207``)
plus iteration-count known answers (``tutorial/step_3/A_getting_statistics.py:43`` = 12 iterations,
``tutorial/step_8/A_visualize_residuals.py:56-58`` = 7 iterations); those are re-asserted in
``tests/test_oracle_reference.py`` / ``oracle/make_golden.py``.  Coefficient-level agreement with
the real qmat beyond round-off is therefore "pinned by properties", not by golden coefficients.
"""
from .qcoeff.collocation import Collocation

Q_GENERATORS = {"Collocation": Collocation, "coll": Collocation}

__all__ = ["Q_GENERATORS", "Collocation"]
