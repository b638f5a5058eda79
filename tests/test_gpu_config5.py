"""BASELINE config 5 in miniature: a union of level-set spheres with its fog volume laid over it, several jittered samples
per pixel.  The reference has no combined pass and VolumeRender has no samples-per-pixel option, so this is an EXTENSION
(SURVEY.md 8d, C5): fog samples use LevelSetRayTracer's jitter rule, and the frame is  fog_film.over(level_set_film)
(Film::RGBA::over, tools/RayTracer.h:252-259).  Its oracle is the restated loop in oracle/vdbrt_oracle.cc; with one sample
per pixel the fog path is still checked against the unmodified reference."""
import numpy as np
import pytest

from openvdb_b200 import api, _abi as abi
from tests import refapi

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 1e-3


def test_film_over_matches_rgba_over(ctx, oracle):
    rng = np.random.default_rng(11)
    top = rng.random((37, 53, 4)).astype(np.float32)
    bot = rng.random((37, 53, 4)).astype(np.float32)
    want = top.copy()
    oracle.film_over(want, bot)
    # the formula itself, in float32 without contraction
    s = bot[..., 3] * (np.float32(1.0) - top[..., 3])
    ref = np.stack([top[..., 3] * top[..., c] + s * bot[..., c] for c in range(3)] + [top[..., 3] + s], axis=2)
    assert np.array_equal(want, ref)
    got = top.copy()
    ctx.film_over(got, bot)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("spp", [1, 4, 16])
def test_fog_supersampling_vs_oracle(ctx, oracle, dfog_small, spp):
    g, og = dfog_small
    W, H = 96, 64
    cam = api.vdb_render_camera(W, H, (10.0, 25.0, 130.0), (0, 0, 0))
    vo = api.vol_opts_default(spp=spp, seed=0)
    vo.primary_step = 0.5
    film = refapi.new_film(W, H)
    ctx.render_volume(g, cam, vo, film)
    want = refapi.new_film(W, H)
    oracle.render_volume(og, cam, vo, want, threads=4)
    assert (want[..., 3] > 0).sum() > 1000
    assert np.array_equal(film[..., 3] > 0, want[..., 3] > 0)
    assert np.allclose(film, want, rtol=RTOL, atol=ATOL)
    print("fog spp %d: %.4f%% pixels not bit-identical" % (spp, 100.0 * float((film != want).any(axis=2).mean())))


def test_config5_overlay_vs_oracle(ctx, oracle, union_dev):
    """level set (16 spp, diffuse) under its own fog volume (16 spp): every pixel within tolerance of the restated loop, the
    level-set layer bit-exact"""
    ls, fog, ols, ofog = union_dev
    W, H = 128, 96
    cam = api.vdb_render_camera(W, H, (40.0, 60.0, 420.0), (0, 0, 0))
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    spp = 16
    f_ls = refapi.new_film(W, H)
    ctx.render_levelset(ls, cam, sh, f_ls, spp=spp, seed=0)
    vo = api.vol_opts_default(spp=spp, seed=0)
    vo.primary_step = 0.5
    f_fog = refapi.new_film(W, H)
    ctx.render_volume(fog, cam, vo, f_fog)
    frame = f_fog.copy()
    ctx.film_over(frame, f_ls)

    o_ls = refapi.new_film(W, H)
    oracle.render_levelset(ols, cam, sh, o_ls, spp=spp, jitter=api.jitter_table(0), threads=4)
    o_fog = refapi.new_film(W, H)
    oracle.render_volume(ofog, cam, vo, o_fog, threads=4)
    want = o_fog.copy()
    oracle.film_over(want, o_ls)
    assert np.array_equal(f_ls, o_ls)
    assert (o_ls[..., 0] > 0).sum() > 1500 and (o_fog[..., 3] > 0).sum() > 1500
    assert np.allclose(f_fog, o_fog, rtol=RTOL, atol=ATOL)
    assert np.allclose(frame, want, rtol=RTOL, atol=ATOL)
    print("C5 overlay: %.4f%% pixels not bit-identical" % (100.0 * float((frame != want).any(axis=2).mean())))


@pytest.fixture(scope="module")
def dfog_small(ctx, oracle):
    ls = ctx.build_sphere(40.0, (3.0, -2.0, 1.0))
    fog = ctx.build_fog(ls)
    og = oracle.open(fog.download())
    yield fog, og
    fog.free(); ls.free()


@pytest.fixture(scope="module")
def union_dev(ctx, oracle):
    rng = np.random.default_rng(20240607)
    s = np.column_stack([rng.uniform(-110, 110, (30, 3)), rng.uniform(10, 40, 30)])
    ls = ctx.build_spheres(s)
    fog = ctx.build_fog(ls)
    ols, ofog = oracle.open(ls.download()), oracle.open(fog.download())
    yield ls, fog, ols, ofog
    fog.free(); ls.free()
