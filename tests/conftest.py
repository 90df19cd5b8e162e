import os
import sys

import pytest

# "virtual" band sets put up to 8 bands (8 contexts, 4 streams each) on one device and synchronise them with spinning flag
# barriers: give the device its maximum of hardware launch queues so that two bands' streams do not share one (must be set
# before CUDA initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The plain-C restatement oracle (oracle/svo_oracle.c), built on demand."""
    from oracle import binding
    return binding.get("orc")


@pytest.fixture(scope="session")
def ref():
    """The reference's own sources compiled through oracle/ref_shim (only where it was built)."""
    from oracle import binding
    if not binding.have_ref():
        if os.path.isdir("/root/reference"):
            binding.build("ref")
        else:
            pytest.skip("oracle/_ref/libsvo_ref.so not built (needs /root/reference)")
    return binding.get("ref")
