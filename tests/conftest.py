import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the unmodified reference node spells numpy.core.records (deprecated alias in numpy 2)
    config.addinivalue_line("filterwarnings", "ignore:numpy.core is deprecated:DeprecationWarning")
