import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


# Code paths written after the round's GPU budget was spent: they are built for sm_100a, their host side is covered by the CPU
# suite and their kernels by the host emulation of the same kernel source (tests/test_kernel_emulation.py), but they have not
# executed on a B200 yet.  Their GPU tests run with MLB_RUN_UNVERIFIED=1 and are skipped - visibly - otherwise, so that the
# parity gate reports on what has been measured on hardware.
UNVERIFIED_ON_HARDWARE = pytest.mark.skipif(os.environ.get("MLB_RUN_UNVERIFIED") != "1",
                                            reason="written after the round-2 GPU budget ran out: not yet run on a B200 (set MLB_RUN_UNVERIFIED=1)")
