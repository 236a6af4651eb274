import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(GOLD, name)) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def gold_pairs():
    return load_golden("pairs.json")


@pytest.fixture(scope="session")
def gold_graphs():
    return load_golden("graphs.json")


@pytest.fixture(scope="session")
def gold_kmer():
    return load_golden("kmer_occurrences.json")


@pytest.fixture(scope="session")
def gold_c1():
    return load_golden("c1.json")


@pytest.fixture(scope="session", params=["pipeline_t1", "pipeline_t2", "pipeline_tricky"])
def gold_pipeline(request):
    d = os.path.join(GOLD, request.param)
    with open(os.path.join(d, "golden.json")) as fh:
        meta = json.load(fh)
    meta["dir"] = d
    return meta
