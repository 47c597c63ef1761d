import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run with -m gpu on the B200 box")
    config.addinivalue_line("markers", "needs_reference: imports the real reference from /root/reference "
                                       "(build container only; skipped elsewhere)")


def pytest_collection_modifyitems(config, items):
    from oracle import ref_import
    if ref_import.reference_available():
        return
    skip = pytest.mark.skip(reason="/root/reference is not present on this machine")
    for item in items:
        if "needs_reference" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dp():
    import json
    import numpy as np
    g = os.path.join(ROOT, "tests", "golden")
    return json.load(open(os.path.join(g, "dp_cases.json"))), np.load(os.path.join(g, "dp_cases.npz"))


@pytest.fixture(scope="session")
def golden_align():
    import json
    import numpy as np
    g = os.path.join(ROOT, "tests", "golden")
    return json.load(open(os.path.join(g, "align_cases.json"))), np.load(os.path.join(g, "align_cases.npz"))


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from whisperx_b200.engine import get_engine
    return get_engine("cuda:0")
