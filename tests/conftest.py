import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "tetra-nerf_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cube_mesh():
    from tetranerf.b200 import synthetic as syn

    return syn.CUBE_VERTICES.copy(), syn.CUBE_CELLS.copy()


@pytest.fixture(scope="session")
def small_mesh():
    from tetranerf.b200 import synthetic as syn

    return syn.delaunay_mesh(3000, seed=0)


@pytest.fixture(scope="session")
def medium_mesh():
    from tetranerf.b200 import synthetic as syn

    return syn.delaunay_mesh(20000, seed=5)
