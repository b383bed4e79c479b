import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_model():
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_model.npz")))


@pytest.fixture(scope="session")
def golden_losses():
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_losses.npz")))


@pytest.fixture(scope="session")
def golden_model_options():
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_model_options.npz")))
