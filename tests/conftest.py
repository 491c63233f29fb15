import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dsm-framework_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_sessionfinish(session, exitstatus):
    """DSMFM_GUARD=1: device buffers carry guard bands (include/dsmfm.h); a run that wrote outside a buffer fails."""
    if os.environ.get("DSMFM_GUARD") and "dsmfm" in sys.modules:
        n = sys.modules["dsmfm"].lib().dsmfm_dbg_guard_violations()
        if n:
            print("\n[dsmfm guard] %d device buffer(s) were overwritten outside their bounds" % n)
            session.exitstatus = 1
