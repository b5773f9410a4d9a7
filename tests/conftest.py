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
