import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref (the compiled reference); skipped when it is absent")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The checkers (oracle port, and the reference when its sources are here) and the product library."""
    import oracle_lib as ol
    if not os.path.exists(os.path.join(ol.ORACLE_DIR, "liboracle.so")) or \
            (os.path.isdir(ol.REFERENCE_SRC) and not ol.have_ref()):
        ol.build_oracle()
    import nbldpc
    if not os.path.exists(nbldpc.LIB_PATH):
        nbldpc.build()
    yield
