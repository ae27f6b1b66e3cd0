import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import _hgt_path  # noqa: E402

_hgt_path.load()

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_NAMES = sorted(f[:-8] for f in os.listdir(GOLDEN_DIR) if f.endswith(".json.gz"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


_cache = {}


def load_golden(name):
    if name not in _cache:
        with gzip.open(os.path.join(GOLDEN_DIR, name + ".json.gz"), "rt") as f:
            _cache[name] = json.load(f)
    return _cache[name]


@pytest.fixture(params=GOLDEN_NAMES)
def golden(request):
    return load_golden(request.param)
