import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # property tests: the same examples on every run unless HYPOTHESIS_PROFILE=explore (or --hypothesis-seed) asks for fresh ones
    try:
        from hypothesis import settings
        settings.register_profile("repeatable", derandomize=True)
        settings.register_profile("explore", derandomize=False)
        settings.load_profile(os.environ.get("HYPOTHESIS_PROFILE", "repeatable"))
    except ImportError:
        pass


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "programs.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_domain():
    with open(os.path.join(ROOT, "tests", "golden", "domain.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_g1():
    with open(os.path.join(ROOT, "tests", "golden", "g1.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.build()
    return binding
