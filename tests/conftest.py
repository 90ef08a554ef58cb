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


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests need a B200: without one they are skipped (with the reason), never run against a fallback --
    the product has none.  A plain `pytest tests` on a CPU box therefore reports them as skipped instead of failing."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    try:
        import nbldpc
        ndev = nbldpc.device_count() if os.path.exists(nbldpc.LIB_PATH) else -1
    except Exception:                                         # library not built yet: the session fixture builds it
        ndev = -1
    if ndev == 0:
        skip = pytest.mark.skip(reason="no CUDA device: the product has no CPU fallback (run `pytest -m gpu` under gpurun)")
        for it in gpu_items:
            it.add_marker(skip)


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Make lost coverage visible: every test skipped for a missing matrix/reference build is counted and named."""
    skipped = terminalreporter.stats.get("skipped", [])
    lost = [r for r in skipped if "not available" in str(r.longrepr) or "oracle/_ref" in str(r.longrepr)]
    if lost:
        terminalreporter.write_line("WARNING: %d test(s) skipped because oracle/_ref (compiled reference + matrices) is missing: "
                                    "run __graft_entry__.build() where /root/reference exists" % len(lost), yellow=True)
        if os.environ.get("NBLDPC_REQUIRE_REF"):
            terminalreporter.write_line("NBLDPC_REQUIRE_REF is set: treating the missing reference build as a failure", red=True)
            terminalreporter._session.exitstatus = 1


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
