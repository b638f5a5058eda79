"""The BASELINE configurations at their FULL sizes against the checkers ("parity holes"):

  * C3: the fog sphere (r = 509, 1024^3 box) at 1920x1080 against the unmodified reference's VolumeRender (threaded), every pixel;
  * C4: the union of 10 000 spheres at 3840x2160 -- the WHOLE film against the unmodified reference's LevelSetRayTracer on the grid
    the reference builds itself (csgUnion of 10 000 createLevelSetSphere), and the per-pixel records (hit, first-hit voxel, t,
    position, normal) of one rank's share of a 4-way split (2.07 M pixels) against the oracle port on the GPU-built buffer;
  * C5: the same union under its own fog volume at 1920x1080 with 16 jittered samples per pixel: the level-set layer of the whole frame
    bit-exact against the oracle port, fog and overlay on one rank's share of a 64-way split (the CPU needs ~10 s for those 32 k pixels);
  * the reference's own VolumeRayIntersector known answers (TestVolumeRayIntersector.cc:30-250) through k_volume_spans;
  * the fog kernel's work counters against the oracle's (they feed B_fog of the roofline).
"""
import os

import numpy as np
import pytest
import torch

from openvdb_b200 import api, _abi as abi
from tests import refapi
from tests.test_gpu_parity import assert_records_equal
from tests.test_reference_kats import VOLUME_KATS

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 1e-3                     # north_star: fog colours within 1e-4 rel / 1e-3 abs (exp is CUDA's on the device)
THREADS = os.cpu_count() or 4
HAVE_REF = os.path.exists(refapi.REF_SO)


def big_memory():
    free, _ = torch.cuda.mem_get_info()
    return free > 60e9


# ---------------------------------------------------------------------------------------------------------------------------
def test_config3_fog_full_1080p(ctx, oracle):
    ls = ctx.build_sphere(509.0)
    fog = ctx.build_fog(ls)
    ls.free()
    W, H = 1920, 1080
    tr, look = (0.0, 0.0, 3 * 509.0), (0.0, 0.0, 0.0)
    cam = api.vdb_render_camera(W, H, tr, look)
    vo = api.vol_opts_default()
    vo.primary_step = 0.5
    film = refapi.new_film(W, H)
    ctx.render_volume(fog, cam, vo, film)
    want = refapi.new_film(W, H)
    if HAVE_REF:
        ref = refapi.Ref()
        ref.set_threads(THREADS)
        rls = ref.sphere(509.0)
        rfog = ref.fog_from_levelset(rls)
        ref.free(rls)
        rvo = ref.vol_defaults()
        rvo.primary_step = 0.5
        ref.render_volume(rfog, refapi.camera_desc(W, H, translation=tr, lookat=look), rvo, want, threaded=True)
        ref.free(rfog)
        ref.set_threads(1)
    else:
        og = oracle.open(fog.download())
        oracle.render_volume(og, cam, vo, want, threads=THREADS)
        oracle.close(og)
    assert (want[..., 3] > 0).sum() > 800000
    assert np.array_equal(film[..., 3] > 0, want[..., 3] > 0)
    assert np.allclose(film, want, rtol=RTOL, atol=ATOL)
    print("C3 1080p: %d of %d pixels not bit-identical, max |diff| %g" % (int((film != want).any(axis=2).sum()), W * H, float(np.abs(film - want).max())))
    fog.free()


def test_fog_counters_match_oracle(ctx, oracle):
    """the numbers that feed B_fog: rays, root / upper / lower probes, primary / shadow samples, shadow rays, alpha>0 pixels.
    The GPU walks the spans lazily (a saturated ray stops walking), so its node probes are a LOWER bound of the oracle's, which
    collects every span of the chord first like the reference; samples and shadow rays are the same set."""
    ls = ctx.build_sphere(100.0)
    fog = ctx.build_fog(ls)
    ls.free()
    og = oracle.open(fog.download())
    W, H = 256, 256
    cam = api.vdb_render_camera(W, H, (0, 0, 300), (0, 0, 0))
    vo = api.vol_opts_default()
    vo.primary_step = 0.5
    c = ctx.count_volume(fog, cam, vo).as_dict()
    film = refapi.new_film(W, H)
    o = oracle.render_volume(og, cam, vo, film, counters=True, threads=1).as_dict()
    for k in ("rays", "primary_samples", "shadow_samples", "shadow_rays", "hits"):
        assert c[k] == o[k], (k, c[k], o[k])
    for k in ("root_probes", "upper_probes", "lower_probes"):
        assert 0 < c[k] <= o[k], (k, c[k], o[k])
    print("fog counters GPU / oracle:", {k: (c[k], o[k]) for k in ("root_probes", "upper_probes", "lower_probes")})
    oracle.close(og)
    fog.free()


# ---------------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c4(ctx):
    if not big_memory():
        pytest.skip("needs ~60 GB of device memory")
    spheres = api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0)
    g = ctx.build_spheres(spheres)
    yield g, spheres
    g.free()


def test_config4_whole_4k_film_against_the_reference(ctx, c4):
    """8 294 400 pixels, bit for bit, against tools::rayTrace of the unmodified reference on the reference's own union"""
    if not HAVE_REF:
        pytest.skip("oracle/_ref not built")
    g, spheres = c4
    W, H = 3840, 2160
    tr, look = (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0)
    cam = api.vdb_render_camera(W, H, tr, look)
    film = refapi.new_film(W, H)
    ctx.render_levelset(g, cam, api.make_shader(abi.SHADER_DIFFUSE), film)
    ref = refapi.Ref()
    rg = ref.spheres_union_mt(spheres, THREADS)
    ref.set_threads(THREADS)
    want = refapi.new_film(W, H)
    ref.render_levelset(rg, refapi.camera_desc(W, H, translation=tr, lookat=look), refapi.shader(abi.SHADER_DIFFUSE), want, threaded=True)
    ref.set_threads(1)
    ref.free(rg)
    assert int((want[..., :3].sum(axis=2) > 0).sum()) > 7000000
    bad = int((film != want).any(axis=2).sum())
    assert bad == 0, "%d of %d pixels differ from the reference's frame" % (bad, W * H)


def test_config4_records_of_a_quarter_of_the_frame(ctx, oracle, c4):
    """rank 1 of a 4-way split: 2.07 M pixels with their hit flag, first-hit voxel, t, position and normal against the oracle port"""
    g, _ = c4
    W, H = 3840, 2160
    cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))
    part = api.partition(1, 4, 64, 60)
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    film = refapi.new_film(W, H, (0.25, 0.5, 0.75, 1.0))
    aux = refapi.AuxArrays(W, H)
    ctx.render_levelset(g, cam, sh, film, aux=aux.pod(), opts=ctx.ls_opts(part=part))
    og = oracle.open(g.download())
    want = refapi.new_film(W, H, (0.25, 0.5, 0.75, 1.0))
    oaux, _ = oracle.render_levelset(og, cam, sh, want, part=part, aux=True, threads=THREADS)
    oracle.close(og)
    assert 1500000 < int(oaux.hit.sum()) < 2073600
    assert_records_equal(aux, oaux)
    assert np.array_equal(film, want)


def test_config5_1080p_16spp_overlay(ctx, oracle, c4):
    if not big_memory():
        pytest.skip("needs ~60 GB of device memory")
    g, _ = c4
    fog = ctx.build_fog(g)
    W, H, spp = 1920, 1080, 16
    cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    # the level-set layer: whole frame, 33 M rays, bit-exact
    f_ls = refapi.new_film(W, H)
    ctx.render_levelset(g, cam, sh, f_ls, spp=spp, seed=0)
    og = oracle.open(g.download())
    o_ls = refapi.new_film(W, H)
    oracle.render_levelset(og, cam, sh, o_ls, spp=spp, jitter=api.jitter_table(0), threads=THREADS)
    oracle.close(og)
    assert (o_ls[..., 0] > 0).sum() > 1500000
    assert np.array_equal(f_ls, o_ls)
    # fog and overlay: one rank's share of a 64-way split (32 400 pixels x 16 samples, each with its shadow rays through 10 000 spheres)
    part = api.partition(37, 64, 64, 60)
    vo = api.vol_opts_default(spp=spp, seed=0)
    vo.primary_step = 0.5
    vo.part = part
    f_fog = refapi.new_film(W, H, (0, 0, 0, 0))
    ctx.render_volume(fog, cam, vo, f_fog)
    ofog = oracle.open(fog.download())
    o_fog = refapi.new_film(W, H, (0, 0, 0, 0))
    oracle.render_volume(ofog, cam, vo, o_fog, threads=THREADS)
    oracle.close(ofog)
    assert (o_fog[..., 3] > 0).sum() > 10000
    assert np.array_equal(f_fog[..., 3] > 0, o_fog[..., 3] > 0)
    assert np.allclose(f_fog, o_fog, rtol=RTOL, atol=ATOL)
    frame, want = f_fog.copy(), o_fog.copy()
    ctx.film_over(frame, f_ls)
    oracle.film_over(want, o_ls)
    assert np.allclose(frame, want, rtol=RTOL, atol=ATOL)
    print("C5 1080p: fog share %d pixels, %d not bit-identical" % (int((o_fog[..., 3] > 0).sum()), int((f_fog != o_fog).any(axis=2).sum())))
    fog.free()


# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", range(len(VOLUME_KATS)))
def test_reference_volume_kats_through_the_kernel(ctx, ref, case):
    """TestVolumeRayIntersector.cc:37-220: one leaf, two leaves, adjacent leaves merging, a gap, an appended active tile, a reversed ray"""
    voxels, boxes, eye, d, want = VOLUME_KATS[case]
    rg = ref.custom(0.0, abi.GRID_CLASS_FOG_VOLUME, 1.0, voxels=voxels, boxes=boxes)
    g = ctx.upload(ref.nanovdb(rg))
    rays = refapi.make_rays([eye], [d])
    spans, counts = ctx.volume_spans(g, rays, space=abi.SPACE_INDEX)
    assert counts[0] == len(want)
    for k, (a, b) in enumerate(want):
        assert abs(spans[0, k, 0] - a) < 1e-6 and abs(spans[0, k, 1] - b) < 1e-6
    rs, rc = ref.volume_spans(rg, rays, space=abi.SPACE_INDEX)
    assert np.array_equal(counts, rc) and np.array_equal(spans, rs)
    g.free()


def test_reference_volume_kat_trevor_through_the_kernel(ctx, ref):
    """TestVolumeRayIntersector.cc:221-250: the ray enters the node bbox but no leaf / tile is active on its path"""
    rg = ref.custom(0.0, abi.GRID_CLASS_FOG_VOLUME, 1.0, voxels=[((0, 0, 0), 1.0), ((20, 20, 0), 1.0)])
    g = ctx.upload(ref.nanovdb(rg))
    rays = refapi.make_rays([(12.5, 4.5, 10.0)], [(0.0, 0.0, -1.0)])
    spans, counts = ctx.volume_spans(g, rays, space=abi.SPACE_INDEX)
    assert counts[0] == 0
    rs, rc = ref.volume_spans(rg, rays, space=abi.SPACE_INDEX)
    assert np.array_equal(counts, rc)
    g.free()
