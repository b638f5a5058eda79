"""Shared fixtures.  Checkers (oracle/) are test infrastructure; the product is openvdb_b200/libvdbrt.so."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests import refapi  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref():
    if not os.path.exists(refapi.REF_SO):
        pytest.skip("oracle/_ref/libvdbref.so not built (needs /root/reference)")
    return refapi.Ref()


@pytest.fixture(scope="session")
def oracle():
    return refapi.Oracle()


@pytest.fixture(scope="session")
def ctx():
    from openvdb_b200 import api
    c = api.Context(0)
    yield c
    c.close()


class GridSet:
    """one synthetic grid, in all three worlds: reference handle, NanoVDB bytes, oracle handle"""

    def __init__(self, ref, oracle, handle):
        self.ref_handle = handle
        self.buf = ref.nanovdb(handle)
        self.oracle_handle = oracle.open(self.buf)


@pytest.fixture(scope="session")
def sphere100(ref, oracle):
    """BASELINE config 1 grid: createLevelSetSphere<FloatGrid>(100, 0, 1, 3)"""
    return GridSet(ref, oracle, ref.sphere(100.0))


@pytest.fixture(scope="session")
def fog100(ref, oracle, sphere100):
    return GridSet(ref, oracle, ref.fog_from_levelset(sphere100.ref_handle))


@pytest.fixture(scope="session")
def sphere_small(ref, oracle):
    """dx = 0.5 sphere off the origin (TestLevelSetRayIntersector's first case: r=5 at (20,0,0), dx=0.5, hw=2)"""
    return GridSet(ref, oracle, ref.sphere(5.0, (20.0, 0.0, 0.0), 0.5, 2.0))


@pytest.fixture(scope="session")
def torus_small(ref, oracle):
    return GridSet(ref, oracle, ref.torus(60.0, 25.0))


@pytest.fixture(scope="session")
def union_small(ref, oracle):
    rng = np.random.default_rng(20240607)
    s = np.column_stack([rng.uniform(-150, 150, (24, 3)), rng.uniform(10, 40, 24)])
    return GridSet(ref, oracle, ref.spheres_union(s))
