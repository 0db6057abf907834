import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    from tests import cases
    return np.load(os.path.join(cases.GOLDEN_DIR, "decks.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_ccl():
    import numpy as np
    from tests import cases
    return np.load(os.path.join(cases.GOLDEN_DIR, "ccl.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_generic():
    import numpy as np
    from tests import cases
    return np.load(os.path.join(cases.GOLDEN_DIR, "generic.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_reuse():
    import numpy as np
    from tests import cases
    return np.load(os.path.join(cases.GOLDEN_DIR, "reuse.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_parosol():
    import numpy as np
    from tests import cases
    return np.load(os.path.join(cases.GOLDEN_DIR, "parosol.npz"), allow_pickle=False)
