"""VolumeRender as a wavefront (csrc/vdbrt_fog.cuh: primary rays -> records of the dense samples -> shadow rays -> pixels) gives the
frame of the one-loop kernel BIT FOR BIT -- same samples, same operands, same order of every floating-point operation -- whatever
the batch size and the record budget (a tile that runs out of record space is re-rendered by the one-loop kernel), for one and
for several samples per pixel, whole and partitioned frames; and both agree with the oracle within the fog tolerance."""
import numpy as np
import pytest

from openvdb_b200 import api, _abi as abi
from tests import refapi

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 1e-3


@pytest.fixture()
def tuned(ctx):
    yield ctx
    ctx.set_tuning(fog_wave=1, fog_refill=8, fog_rec_per_ray=12, fog_cap_mb=4096)


@pytest.fixture(scope="module")
def fog_union(ctx, oracle):
    rng = np.random.default_rng(20240607)
    s = np.column_stack([rng.uniform(-110, 110, (30, 3)), rng.uniform(10, 40, 30)])
    ls = ctx.build_spheres(s)
    fog = ctx.build_fog(ls)
    ls.free()
    og = oracle.open(fog.download())
    yield fog, og
    fog.free()


def render(ctx, fog, cam, W, H, spp=1, part=None, init=(0.5, 0.25, 0.125, 0.75)):
    vo = api.vol_opts_default(spp=spp, seed=2)
    vo.primary_step = 0.5
    if part is not None:
        vo.part = part
    film = refapi.new_film(W, H, init)
    ctx.render_volume(fog, cam, vo, film)
    return film, vo


@pytest.mark.parametrize("spp", [1, 3])
def test_wavefront_equals_one_loop_kernel_and_oracle(tuned, oracle, fog_union, spp):
    ctx = tuned
    fog, og = fog_union
    W, H = 203, 117                       # edge tiles with slots outside the film
    cam = api.vdb_render_camera(W, H, (40.0, 60.0, 420.0), (0, 0, 0))
    ctx.set_tuning(fog_wave=0)
    base, vo = render(ctx, fog, cam, W, H, spp)
    assert ctx.last_kernel_ms()[1] == 1
    want = refapi.new_film(W, H, (0.5, 0.25, 0.125, 0.75))
    oracle.render_volume(og, cam, vo, want, threads=8)
    assert (want[..., 3] > 0).sum() > 5000
    assert np.array_equal(base[..., 3] > 0, want[..., 3] > 0) and np.allclose(base, want, rtol=RTOL, atol=ATOL)
    # (records per ray, MB per batch): plenty / several batches / starved (most tiles fall back) / one record for the whole batch
    for per_ray, cap_mb, refill in ((12, 4096, 8), (12, 1, 32), (2, 4096, 1), (1, 1, 8), (12, 4096, 16)):
        ctx.set_tuning(fog_wave=1, fog_rec_per_ray=per_ray, fog_cap_mb=cap_mb, fog_refill=refill)
        film, _ = render(ctx, fog, cam, W, H, spp)
        assert ctx.last_kernel_ms()[1] >= 4
        assert np.array_equal(film, base), (per_ray, cap_mb, refill)
    # three ranks, one after the other, into one film
    ctx.set_tuning(fog_wave=1, fog_refill=8, fog_rec_per_ray=12, fog_cap_mb=4096)
    film = refapi.new_film(W, H, (0.5, 0.25, 0.125, 0.75))
    for r in range(3):
        vo = api.vol_opts_default(spp=spp, seed=2)
        vo.primary_step = 0.5
        vo.part = api.partition(r, 3, 16, 8)
        ctx.render_volume(fog, cam, vo, film)
    assert np.array_equal(film, base)


def test_wavefront_thin_fog_many_records_per_ray(tuned, oracle):
    """low density -> the primary ray does not saturate -> hundreds of dense samples (records) per ray: batches shrink, tiles fall back"""
    ctx = tuned
    ls = ctx.build_sphere(60.0)
    fog = ctx.build_fog(ls)
    ls.free()
    og = oracle.open(fog.download())
    W, H = 96, 64
    cam = api.vdb_render_camera(W, H, (0.0, 0.0, 180.0), (0, 0, 0))
    vo = api.vol_opts_default()
    vo.primary_step = 0.5
    vo.absorption = abi.vec3((0.001, 0.002, 0.003))
    vo.scattering = abi.vec3((0.004, 0.003, 0.002))
    vo.light_dir = abi.vec3(tuple(np.array((0.0, 1.0, 0.2)) / np.linalg.norm((0.0, 1.0, 0.2))))
    want = refapi.new_film(W, H)
    oracle.render_volume(og, cam, vo, want, threads=8)
    frames = []
    for wave, per_ray in ((0, 12), (1, 12), (1, 400)):
        ctx.set_tuning(fog_wave=wave, fog_rec_per_ray=per_ray)
        film = refapi.new_film(W, H)
        ctx.render_volume(fog, cam, vo, film)
        frames.append(film)
    assert np.array_equal(frames[0], frames[1]) and np.array_equal(frames[0], frames[2])
    assert np.array_equal(frames[0][..., 3] > 0, want[..., 3] > 0) and np.allclose(frames[0], want, rtol=RTOL, atol=ATOL)
    oracle.close(og)
    fog.free()
