"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests of the widened rows (AETHER post, smoke march, viewshed, LBVH: SURVEY section 8f) run after the core hot-path GPU tests, so
    that under `-x` a failure in a newer row cannot hide the state of the section 8a-e parity suite."""
    late = {"test_aether.py", "test_smoke.py", "test_viewshed.py", "test_lbvh.py", "test_wavefront.py"}
    items.sort(key=lambda it: 1 if (it.get_closest_marker("gpu") and Path(str(it.fspath)).name in late) else 0)
