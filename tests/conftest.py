import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
GOLDEN_MODELS = ["getting_started", "coffeecup", "cutting_disc", "plate_ffe", "plate_mor", "plate_ffe_var",
                 "plate_mor_var", "kuhn_cuboid", "tdep_rod", "assembly", "lp_tdep", "steel_sheet",
                 "plate_ffe_stat", "plate_ffe_statc", "plate_mor_stat"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden():
    from thermca_b200.spec import load_spec
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_spec(os.path.join(GOLDEN, name + ".npz"))
        return cache[name]
    return get
