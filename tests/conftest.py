import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def vldb():
    """Config C1 input: the reference's data/vldb_2025.parquet embedding column (496 x 4096 f32)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "vldb_2025_embeddings.npz"))
    return np.ascontiguousarray(z["embedding"], dtype=np.float32)
