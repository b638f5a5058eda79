"""Host films: a pinned film is written by the kernels in place (no device->host copy), a pageable one goes through the
staging buffer.  Both must hold exactly the pixels of the oracle, with a uniform or an arbitrary old film, for a whole
frame and for one rank's share of it (the other ranks' pixels stay untouched)."""
import numpy as np
import pytest

from openvdb_b200 import api, _abi as abi
from tests import refapi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene(ctx, oracle):
    ls = ctx.build_sphere(40.0, (3.0, -2.0, 1.0))
    fog = ctx.build_fog(ls)
    yield ls, fog, oracle.open(ls.download()), oracle.open(fog.download())
    fog.free(); ls.free()


def test_levelset_pinned_and_pageable_films(ctx, oracle, scene):
    ls, _, ols, _ = scene
    W, H = 224, 160
    cam = api.vdb_render_camera(W, H, (30.0, 20.0, 140.0), (0, 0, 0))
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    rng = np.random.default_rng(9)
    old = rng.random((H, W, 4)).astype(np.float32)
    want = old.copy()
    oracle.render_levelset(ols, cam, sh, want, threads=4)
    pinned = api.PinnedArray((H, W, 4), np.float32)
    pinned.array[...] = old
    ctx.render_levelset(ls, cam, sh, pinned.array)                    # arbitrary old film, pinned: staged input, in-place output
    assert np.array_equal(pinned.array, want)
    pageable = old.copy()
    ctx.render_levelset(ls, cam, sh, pageable)
    assert np.array_equal(pageable, want)
    # uniform old film: nothing is copied to the device at all
    bg = (0.25, 0.5, 0.75, 1.0)
    want_u = refapi.new_film(W, H, bg)
    oracle.render_levelset(ols, cam, sh, want_u, threads=4)
    pinned.array[...] = bg
    ctx.render_levelset(ls, cam, sh, pinned.array, uniform_bg=True, bg=bg)
    assert np.array_equal(pinned.array, want_u)
    # one rank of three: only its tiles change
    for film in (pinned.array, old.copy()):
        film[...] = old
        ctx.render_levelset(ls, cam, sh, film, part=api.partition(1, 3, 32, 32))
        mine = np.zeros((H, W), bool)
        tiles_x = (W + 31) // 32
        for ty in range((H + 31) // 32):
            for tx in range(tiles_x):
                if (ty * tiles_x + tx) % 3 == 1:
                    mine[ty * 32:(ty + 1) * 32, tx * 32:(tx + 1) * 32] = True
        assert np.array_equal(film[mine], want[mine]) and np.array_equal(film[~mine], old[~mine])


def test_volume_pinned_and_pageable_films(ctx, oracle, scene):
    _, fog, _, ofog = scene
    W, H = 128, 96
    cam = api.vdb_render_camera(W, H, (30.0, 20.0, 140.0), (0, 0, 0))
    vo = api.vol_opts_default()
    vo.primary_step = 0.5
    want = refapi.new_film(W, H)
    oracle.render_volume(ofog, cam, vo, want, threads=4)
    pinned = api.PinnedArray((H, W, 4), np.float32)
    pinned.array[...] = 7.0
    ctx.render_volume(fog, cam, vo, pinned.array)
    pageable = np.full((H, W, 4), 7.0, np.float32)
    ctx.render_volume(fog, cam, vo, pageable)
    assert np.array_equal(pinned.array, pageable)
    assert np.array_equal(pageable[..., 3] > 0, want[..., 3] > 0) and np.allclose(pageable, want, rtol=1e-4, atol=1e-3)
