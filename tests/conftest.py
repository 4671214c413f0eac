import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "eao-fusion_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def tex640():
    from eaof import synth
    return synth.base_texture(640, 480)


@pytest.fixture(scope="session")
def frames640(tex640):
    from eaof import synth
    return synth.make_frames(6, 640, 480, tex=tex640)
