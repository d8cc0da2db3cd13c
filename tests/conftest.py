import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

SHAPES = [(9, 9, 6), (6, 6, 4), (12, 12, 7), (16, 16, 8), (6, 6, 3), (5, 5, 2)]
# round 2: every other square size boardConfig accepts, with smaller reference-generated fixtures
EXTRA_SHAPES = [(4, 4, 4), (7, 7, 5), (8, 8, 5), (10, 10, 6), (11, 11, 7), (13, 13, 9), (14, 14, 6), (15, 15, 8)]
ALL_SHAPES = SHAPES + EXTRA_SHAPES


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # build what a fresh checkout lacks (built artefacts are git-ignored): the CUDA library (nvcc cross-compiles
    # without a GPU) and the CPU oracle; the host simulator builds itself on first use
    lib = os.path.join(ROOT, "element-crush-gym_b200", "lib", "libecg.so")
    if not os.path.exists(lib):
        import importlib
        importlib.import_module("element-crush-gym_b200.build").build()
    from oracle import oracle as _o
    _o.build()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
