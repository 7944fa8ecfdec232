import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


@pytest.fixture(scope="session")
def weights(golden):
    return dict(golden("weights"))


def sub_sd(sd, prefix):
    """Strip `prefix` from a state-dict style mapping."""
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


@pytest.fixture(autouse=True, params=["tensor", "ffma"])
def conv_mode(request):
    """Every parity test runs in both convolution arithmetics (include/pmctf_b200.h PMCTF_CONV_*): the oracle and,
    for GPU tests, the CUDA library are switched together and must agree bit for bit in each."""
    from oracle import oracle as orc
    orc.set_conv_mode(request.param)
    gpu = request.node.get_closest_marker("gpu") is not None
    if gpu:
        import learned_pmctf_b200 as pkg
        pkg.ops.set_conv_mode(request.param)
    yield request.param
    orc.set_conv_mode("ffma")
    if gpu:
        import learned_pmctf_b200 as pkg
        assert pkg._native.lib().pmctf_tc_error_flag() == 0, "a tensor-core kernel timed out waiting for an MMA"
