"""pytest configuration: markers and import paths.

`-m "not gpu"` covers the oracle against the reference's golden vectors, the host
logic and the exported C ABI (no GPU needed); `-m gpu` holds the parity tests
proper, which call the CUDA library through its C ABI on a real B200.
"""
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(REPO, "diff-volume-renderer_b200", "python"), os.path.join(REPO, "oracle"),
          os.path.join(REPO, "tests"), REPO):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    import oracle
    oracle.build_oracle()
