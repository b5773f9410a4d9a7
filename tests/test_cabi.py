"""The C ABI: libsdcb200.so loads without a GPU and exports exactly what include/sdc_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sdc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdcb200_\w+)\s*\(", src)))


def test_library_builds_and_exports_header():
    from pysdc_b200.build import build
    from pysdc_b200.backend import load_library

    build()
    lib = load_library()
    names = _declared()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/sdc_b200.h but not exported"
    assert sorted(lib._sdc_signatures) == names, "backend.py binds a different set of functions than the header declares"
    assert lib.sdcb200_version() >= 100
    # pure host helpers can be called without a GPU
    assert lib.sdcb200_pitch(511) == 512 and lib.sdcb200_pitch(2048) == 2048
    assert lib.sdcb200_volume(3, 511) == 512**3 and lib.sdcb200_guard(3, 511) == 512**2
    assert lib.sdcb200_guard(1, 1023) == 16


def test_no_cpu_fallback_without_gpu():
    import torch
    from pysdc_b200 import backend
    from pysdc_b200.errors import BackendError

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    old = backend._backend
    backend.set_backend(None)
    try:
        with pytest.raises(BackendError):
            backend.get_backend()
        from pysdc_b200.problems import heatNd_unforced
        with pytest.raises(BackendError):
            heatNd_unforced(nvars=(31, 31), nu=0.1, freq=(2, 2), bc="dirichlet-zero", solver_type="CG")
    finally:
        backend.set_backend(old)


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pysdc_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "sdc_oracle" not in src and "qmat_shim" not in src and "import oracle" not in src, f
