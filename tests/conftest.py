import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def golden_cases(prefix):
    # (softpool_feat.npz is the SoftPoolFeat module fixture: other layout, its own test)
    return sorted(f[len(prefix) + 1:-4] for f in os.listdir(GOLDEN)
                  if f.startswith(prefix + "_") and f.endswith(".npz") and f != "softpool_feat.npz")
