import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 93923590            # leven/default.cfg:8


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def oracle_mod(built):
    from oracle import oracle as O
    return O


@pytest.fixture(scope="session")
def world(oracle_mod):
    w = oracle_mod.World(seed=SEED, default_material=0, voxels_per_chunk=64)
    yield w
    w.close()


@pytest.fixture(scope="session")
def surface_cy(world):
    """config 1: the chunk above the world origin that contains the surface"""
    return int(900 * world.terrain(0.0, 0.0) // 64)


@pytest.fixture(scope="session")
def golden_keys():
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "octree_keys.npz")))


@pytest.fixture(scope="session")
def lc(built):
    """the product's host mirror, initialised on cuda:0 (GPU tests only)"""
    import leven_b200.compute as lc
    rc = lc.Compute_Initialise(SEED, 0, 2)
    if rc < 0:
        pytest.fail(f"Compute_Initialise failed: {lc.GetCLErrorString(rc)} {lc.last_cuda_error()} "
                    "(GPU tests need the CUDA path; there is no fallback)")
    return lc
