import os
import sys
import warnings

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
warnings.filterwarnings("ignore", category=DeprecationWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # The generated-input tests draw the same examples on every run (a regular test run is reproducible);
    # SSHG_HYP_SCALE > 1 (the soak runs, tests/test_gpu_hypothesis.py) draws fresh ones.
    try:
        from hypothesis import settings
        settings.register_profile("reproducible", derandomize=True)
        settings.register_profile("soak", derandomize=False)
        settings.load_profile("soak" if int(os.environ.get("SSHG_HYP_SCALE", "1")) > 1 else "reproducible")
    except ImportError:
        pass


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def libopt():
    """Test / tuning switches of the library's process-wide context for the rest of the test -- libopt(k3=0,
    k2_sparse=1) -- restored afterwards (sshg_ctx_set_option; the library reads no environment variables)."""
    import contextlib
    from shgyield_b200 import _lib
    with contextlib.ExitStack() as stack:
        yield lambda **kv: stack.enter_context(_lib.context().options(**kv))
