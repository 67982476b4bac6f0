import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ctx():
    import hite_b200
    c = hite_b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def golden_v5():
    import gzip
    import json
    with gzip.open(os.path.join(ROOT, "tests", "golden", "judge_v5.json.gz")) as f:
        return json.loads(f.read())


@pytest.fixture(scope="session")
def golden_v6():
    import gzip
    import json
    with gzip.open(os.path.join(ROOT, "tests", "golden", "judge_v6.json.gz")) as f:
        return json.loads(f.read())


@pytest.fixture(scope="session")
def golden_v9():
    import gzip
    import json
    with gzip.open(os.path.join(ROOT, "tests", "golden", "judge_v9.json.gz")) as f:
        return json.loads(f.read())
