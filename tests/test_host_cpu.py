"""CPU suite for the HOST logic (datatypes, problems, sweepers, controller) against the reference's golden fixtures.
The kernel library is replaced by the numpy test double in tests/fake_backend.py; the real kernels are exercised by
tests/test_gpu_parity.py (-m gpu) with the very same checks."""
import numpy as np
import pytest

import parity_cases as pc
from conftest import golden_names


@pytest.fixture(autouse=True)
def numpy_backend():
    from fake_backend import NumpyBackend
    from pysdc_b200 import backend

    old = backend._backend
    backend.set_backend(NumpyBackend())
    yield
    backend.set_backend(old)


@pytest.mark.parametrize("name", golden_names("op_"))
def test_operator(name):
    pc.check_operator(name)


@pytest.mark.parametrize("name", golden_names("sweep_"))
def test_sweep_dump(name):
    pc.check_sweep_dump(name)


@pytest.mark.parametrize("name", [n for n in golden_names("run_") if "255" not in n and "63_K4" not in n])
def test_run(name):
    # 1-D grids are badly conditioned (kappa ~ 1e4): CG loses orthogonality and its iteration count depends on
    # rounding details at the 10 % level; SDC iteration counts and the solution are unaffected
    pc.check_run(name, count_slack=None if "heat1d" in name else 0.02)
