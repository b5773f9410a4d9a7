import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")
    for name in ("base", "mpi4py", "cupy", "slow", "benchmark"):  # marks of the reference's own tests (test_reference_suite.py)
        config.addinivalue_line("markers", f"{name}: mark used by the reference's test-suite")


def pytest_collection_modifyitems(config, items):
    """GPU tests skip (instead of erroring) where there is no CUDA device or no built library: a plain ``pytest`` on a
    CPU box runs the CPU suite and reports the rest as skipped."""
    try:
        import torch

        have = torch.cuda.is_available()
    except Exception:
        have = False
    have = have and os.path.exists(os.path.join(ROOT, "pysdc_b200", "lib", "libsdcb200.so"))
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and pysdc_b200/lib/libsdcb200.so")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def reference_paths():
    """sys.path entries for the UNMODIFIED reference (+ the qmat stand-in): /root/reference in the build container, the
    shipped copy oracle/_ref (oracle/build_ref.py) on the GPU box; None when neither is there."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref

    return build_ref.reference_paths()


def load_golden(name):
    """Fixture written by oracle/make_golden.py from the unmodified reference: (spec dict, arrays)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    spec = json.loads(str(z["spec"]))
    return spec, {k: z[k] for k in z.files if k != "spec"}


def golden_names(prefix):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith(".npz"))


@pytest.fixture(scope="session")
def oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import sdc_oracle

    return sdc_oracle


def free_port():
    """A TCP port that is free right now on 127.0.0.1 (rendezvous of the multi-process gloo tests)."""
    import socket

    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]
