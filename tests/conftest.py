import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle import loader

    return loader.api("port")


@pytest.fixture(scope="session")
def ref():
    """The reference's own code (oracle/_ref/libsrsref.so).  Built here when /root/reference exists; on the GPU box
    the prebuilt file travels with the snapshot.  Tests that need it skip when neither is available."""
    from oracle import loader

    if not loader.have_ref():
        try:
            loader.build()
        except Exception:
            pass
    if not loader.have_ref():
        pytest.skip("oracle/_ref/libsrsref.so not available (reference tree absent and no prebuilt library)")
    return loader.api("ref")
