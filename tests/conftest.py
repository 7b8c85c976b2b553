import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture
def record():
    """record(name, **values): appends a JSON line to gpurun_out/test_metrics.jsonl (measured
    distances of the parity tests, read back after a GPU run; a no-op without that directory)."""
    import json

    def _rec(name, **values):
        d = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(d):
            with open(os.path.join(d, "test_metrics.jsonl"), "a") as f:
                f.write(json.dumps({"test": name, **values}) + "\n")
    return _rec
