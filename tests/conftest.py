import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def tekken_json():
    from tekken_rs_b200 import assets
    return assets.ensure_tekken_json()


@pytest.fixture(scope="session")
def oracle(tekken_json):
    from oracle import tekken_oracle as TO
    return TO.OracleTekkenizer.from_file(tekken_json)


@pytest.fixture(scope="session")
def goldens():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json"), encoding="utf-8"))


@pytest.fixture(scope="session")
def fixtures():
    import json
    return json.load(open(os.path.join(ROOT, "tests", "golden", "engine_fixtures.json"), encoding="utf-8"))


@pytest.fixture(scope="session")
def gpu_tok(tekken_json):
    """The CUDA tokenizer on device 0.  Fails (does not skip) if the library or device is missing:
    a GPU test must never pass on a fallback."""
    from tekken_rs_b200 import Tekkenizer
    return Tekkenizer.from_file(tekken_json, device=0)


@pytest.fixture(scope="session")
def host_tok(tekken_json):
    """Host-only handle (device -1): accessors and validation, no compute."""
    from tekken_rs_b200 import Tekkenizer
    return Tekkenizer.from_file(tekken_json, device=-1)
