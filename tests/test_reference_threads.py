"""The threaded reference arm (oracle/_ref through the TBB shim) must be race-free at any thread count: every worker's
body copy registers its ValueAccessor in the tree's accessor registry (tree/Tree.h:1081-1082,1423-1450), a
tbb::concurrent_hash_map.  Round 1's shim had an unlocked map there and crashed at 192 host threads."""
import os

import numpy as np
import pytest

from openvdb_b200 import _abi as abi
from tests import refapi

pytestmark = pytest.mark.skipif(not os.path.exists(refapi.REF_SO), reason="oracle/_ref not built")


def test_threaded_render_256_threads_20_times():
    ref = refapi.Ref()
    g = ref.sphere(40.0)
    W, H = 256, 256                      # 256 rows: one chunk per thread at 256 threads
    d = refapi.camera_desc(W, H, translation=(0, 0, 120), lookat=(0, 0, 0))
    sh = refapi.shader(abi.SHADER_DIFFUSE)
    ref.set_threads(1)
    serial = refapi.new_film(W, H)
    ref.render_levelset(g, d, sh, serial, threaded=False)
    try:
        ref.set_threads(256)
        for _ in range(20):
            film = refapi.new_film(W, H)
            ref.render_levelset(g, d, sh, film, threaded=True)
            assert np.array_equal(film, serial)
        fog = ref.fog_from_levelset(g)
        vo = ref.vol_defaults()
        f0 = refapi.new_film(W, H)
        ref.set_threads(1)
        ref.render_volume(fog, d, vo, f0, threaded=False)
        ref.set_threads(256)
        for _ in range(5):
            f1 = refapi.new_film(W, H)
            ref.render_volume(fog, d, vo, f1, threaded=True)
            assert np.array_equal(f0, f1)
    finally:
        ref.set_threads(1)
