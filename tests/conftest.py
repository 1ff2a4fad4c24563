import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "ref: needs the reference tree mounted at /root/reference")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/src/spdepy")
    skip_ref = pytest.mark.skip(reason="reference tree not mounted")
    for item in items:
        if "ref" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
