import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
# tests/test_gpu_shards_local.py runs several shards of one colony in ONE process: a shard's barrier kernel waits for kernels of the
# other shards' streams, so those streams must not share a hardware work queue (default: 8 queues; the suite creates more streams than
# that before it gets there).  Must be set before the CUDA context exists.  One process per GPU — the product layout — is not affected.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def meshes():
    return dict(np.load(os.path.join(GOLDEN, "meshes.npz")))


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(GOLDEN, "ref_kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ref_acs():
    return dict(np.load(os.path.join(GOLDEN, "ref_acs.npz")))


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Built on demand."""
    from oracle import oracle as O
    O.lib()
    return O


C1_POINTS = [(1.600931, y, z) for y in (-0.259319, -0.074319, 0.085681) for z in (1.224003, 1.399003)]
C1_NAN_PAIR = ((1.59, -0.27, 1.22), (1.59, 0.12, 1.22))
