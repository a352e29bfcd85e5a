import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def pkg():
    """The product package (directory name has hyphens, so it is imported through importlib / the rtd3_b200 shim)."""
    return importlib.import_module("rtd3_b200")


@pytest.fixture(scope="session")
def env_golden():
    return load_golden("env_golden.npz")


@pytest.fixture(scope="session")
def replay_golden():
    return load_golden("replay_golden.npz")


@pytest.fixture(scope="session")
def td3_golden():
    return load_golden("td3_golden.npz")


@pytest.fixture(scope="session")
def robot_golden():
    return load_golden("robot_golden.npz")


@pytest.fixture(scope="session")
def trace_golden():
    return load_golden("trace_golden.npz")


@pytest.fixture(scope="session")
def loop_golden():
    return load_golden("loop_golden.npz")
