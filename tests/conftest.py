import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as entry  # noqa: E402
import oracle_lib as ol  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def jb():
    mod = entry.load()
    if not os.path.exists(mod.LIB_PATH):
        mod.build()
    return mod


@pytest.fixture(scope="session")
def enc(jb):
    e = jb.Encoder(0)  # raises without a GPU: the product has no CPU fallback
    yield e
    e.close()


@pytest.fixture(scope="session")
def oracle():
    ol.oracle()
    return ol


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "reference_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def fruit():
    return ol.read_ppm(os.path.join(GOLDEN, "fruit.ppm"))


def noise_image(seed, W, H):
    return np.random.default_rng(seed).integers(0, 256, (H, W, 3), dtype=np.uint8)
