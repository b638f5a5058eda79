"""Parity tests proper: the CUDA path (through the C ABI of libvdbrt.so) against the oracle port, the committed golden
fixtures of the reference, and -- when oracle/_ref travelled to the box -- the unmodified reference itself.

Bar (BASELINE.json north_star): hit/miss mask and first-hit voxel bit-exact; hit distance, normal and RGBA within
1e-4 relative / 1e-3 absolute.  In practice the level-set path is bit-exact everywhere (asserted), the fog path is held
to the tolerance because exp() is CUDA's, not glibc's."""
import ctypes as C
import os

import numpy as np
import pytest

from openvdb_b200 import api, _abi as abi
from tests import refapi

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL, ATOL = 1e-4, 1e-3


def gpu_levelset(ctx, dgrid, cam, sh, W, H, bg=(0, 0, 0, 1), **kw):
    film = refapi.new_film(W, H, bg)
    aux = refapi.AuxArrays(W, H)
    pod = aux.pod()
    ctx.render_levelset(dgrid, cam, sh, film, aux=pod, **kw)
    return film, aux


def assert_records_equal(a, b):
    assert np.array_equal(a.hit, b.hit), "hit mask"
    assert np.array_equal(a.ijk, b.ijk), "first-hit voxel"
    for k in ("t_index", "t_world", "xyz", "nml"):
        x, y = getattr(a, k), getattr(b, k)
        assert np.allclose(x, y, rtol=RTOL, atol=ATOL), k
        assert np.array_equal(x, y), k + " (bit-exact)"


@pytest.fixture(scope="module")
def dsphere(ctx, sphere100):
    g = ctx.upload(sphere100.buf)
    yield g
    g.free()


@pytest.fixture(scope="module")
def dfog(ctx, fog100):
    g = ctx.upload(fog100.buf)
    yield g
    g.free()


def test_upload_derives_node_bbox(ctx, oracle, dsphere, sphere100):
    want = oracle.info(sphere100.oracle_handle)
    assert list(dsphere.info.node_bbox) == list(want.node_bbox) == [-104, -104, -104, 103, 103, 103]
    assert list(dsphere.info.index_bbox) == list(want.index_bbox)
    assert dsphere.info.leaf_count == 4025 and dsphere.info.active_voxels == 753990
    assert dsphere.info.grid_class == abi.GRID_CLASS_LEVEL_SET and dsphere.info.background == 3.0


@pytest.mark.parametrize("kind", [abi.SHADER_DIFFUSE, abi.SHADER_NORMAL, abi.SHADER_MATTE, abi.SHADER_POSITION])
def test_levelset_vs_oracle(ctx, oracle, dsphere, sphere100, kind):
    W, H = 320, 200
    cam = api.vdb_render_camera(W, H, (40, 60, 280), (0, 0, 0))
    sh = api.make_shader(kind, (0.9, 0.7, 0.5, 0.8), bbox_min=(-100, -100, -100), inv_dim=(1 / 200.0,) * 3)
    bg = (0.1, 0.2, 0.3, 0.4)
    film, aux = gpu_levelset(ctx, dsphere, cam, sh, W, H, bg)
    ofilm = refapi.new_film(W, H, bg)
    oaux, _ = oracle.render_levelset(sphere100.oracle_handle, cam, sh, ofilm, aux=True, threads=4)
    assert aux.hit.sum() > 10000
    assert_records_equal(aux, oaux)
    assert np.array_equal(film, ofilm)


def test_levelset_config1_full_size(ctx, oracle, dsphere, sphere100):
    """BASELINE config 1 at its full 1024x1024: 606 028 hit pixels and every record equal to the oracle's"""
    W = H = 1024
    cam = api.vdb_render_camera(W, H, (0, 0, 300), (0, 0, 0))
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    film, aux = gpu_levelset(ctx, dsphere, cam, sh, W, H)
    assert int(aux.hit.sum()) == 606028
    ofilm = refapi.new_film(W, H)
    oaux, _ = oracle.render_levelset(sphere100.oracle_handle, cam, sh, ofilm, aux=True, threads=os.cpu_count() or 4)
    assert_records_equal(aux, oaux)
    assert np.array_equal(film, ofilm)
    # no-aux / uniform-background / device-film variants produce the same film
    f2 = refapi.new_film(W, H)
    ctx.render_levelset(dsphere, cam, sh, f2, uniform_bg=True, bg=(0, 0, 0, 1))
    assert np.array_equal(f2, film)


def test_levelset_vs_reference_directly(ctx, ref, dsphere, sphere100):
    W, H = 256, 256
    d = refapi.camera_desc(W, H, translation=(0, 0, 300), lookat=(0, 0, 0))
    cam = api.vdb_render_camera(W, H, (0, 0, 300), (0, 0, 0))
    assert bytes(cam) == bytes(ref.camera_pod(d))
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    film, aux = gpu_levelset(ctx, dsphere, cam, sh, W, H)
    rfilm = refapi.new_film(W, H)
    ref.render_levelset(sphere100.ref_handle, d, sh, rfilm)
    raux, _, mism = ref.levelset_records(sphere100.ref_handle, d)
    assert mism == 0
    assert_records_equal(aux, raux)
    assert np.array_equal(film, rfilm)
    assert int(aux.hit.sum()) == 37896


def test_levelset_supersampling(ctx, oracle, dsphere, sphere100):
    W, H = 160, 96
    cam = api.vdb_render_camera(W, H, (0, 0, 300), (0, 0, 0))
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    for spp, seed in ((2, 0), (5, 3), (16, 0)):
        film = refapi.new_film(W, H)
        ctx.render_levelset(dsphere, cam, sh, film, spp=spp, seed=seed)
        ofilm = refapi.new_film(W, H)
        oracle.render_levelset(sphere100.oracle_handle, cam, sh, ofilm, spp=spp, jitter=api.jitter_table(seed), threads=4)
        assert np.array_equal(film, ofilm), (spp, seed)


def test_levelset_orthographic_iso_and_scaled_grid(ctx, oracle, dsphere, sphere100, sphere_small):
    W, H = 200, 150
    cam = api.orthographic_camera(W, H, (5, -10, 20), (10, 5, 250), 260.0, float(np.float32(1e-3)), api.FLT_MAX)
    for iso in (0.0, 1.25, -2.0):
        film, aux = gpu_levelset(ctx, dsphere, cam, api.make_shader(abi.SHADER_NORMAL), W, H, iso=iso)
        ofilm = refapi.new_film(W, H)
        oaux, _ = oracle.render_levelset(sphere100.oracle_handle, cam, api.make_shader(abi.SHADER_NORMAL), ofilm, iso=iso, aux=True)
        assert_records_equal(aux, oaux)
        assert np.array_equal(film, ofilm)
    g = ctx.upload(sphere_small.buf)
    cam = api.vdb_render_camera(W, H, (2, 3, 30), (20, 0, 0))
    film, aux = gpu_levelset(ctx, g, cam, api.make_shader(), W, H)
    ofilm = refapi.new_film(W, H)
    oaux, _ = oracle.render_levelset(sphere_small.oracle_handle, cam, api.make_shader(), ofilm, aux=True)
    assert aux.hit.sum() > 1000
    assert_records_equal(aux, oaux)
    assert np.array_equal(film, ofilm)
    g.free()


def test_levelset_torus_and_union(ctx, oracle, torus_small, union_small):
    for gs, tr in ((torus_small, (0, 90, 255)), (union_small, (50, 120, 520))):
        g = ctx.upload(gs.buf)
        W, H = 320, 180
        cam = api.vdb_render_camera(W, H, tr, (0, 0, 0))
        film, aux = gpu_levelset(ctx, g, cam, api.make_shader(), W, H)
        ofilm = refapi.new_film(W, H)
        oaux, _ = oracle.render_levelset(gs.oracle_handle, cam, api.make_shader(), ofilm, aux=True, threads=4)
        assert aux.hit.sum() > 3000
        assert_records_equal(aux, oaux)
        assert np.array_equal(film, ofilm)
        g.free()


def test_ragged_film_sizes_and_partitions(ctx, oracle, dsphere, sphere100):
    """film sizes that are not multiples of the 8x4 warp tile / 64x64 macro tile, 1-pixel films, and a 3-way tile partition
    whose union equals the unpartitioned render"""
    sh = api.make_shader()
    for W, H in ((1, 1), (7, 3), (65, 33), (130, 67)):
        cam = api.vdb_render_camera(W, H, (0, 0, 300), (0, 0, 0))
        film, aux = gpu_levelset(ctx, dsphere, cam, sh, W, H)
        ofilm = refapi.new_film(W, H)
        oaux, _ = oracle.render_levelset(sphere100.oracle_handle, cam, sh, ofilm, aux=True)
        assert_records_equal(aux, oaux)
        assert np.array_equal(film, ofilm)
    W, H = 200, 120
    cam = api.vdb_render_camera(W, H, (0, 0, 300), (0, 0, 0))
    whole = refapi.new_film(W, H, (0.5, 0.5, 0.5, 0.5))
    ctx.render_levelset(dsphere, cam, sh, whole)
    acc = refapi.new_film(W, H, (0.5, 0.5, 0.5, 0.5))
    for r in range(3):
        before = acc.copy()
        ctx.render_levelset(dsphere, cam, sh, acc, part=api.partition(r, 3, 32, 16))
        opart = before.copy()
        oracle.render_levelset(sphere100.oracle_handle, cam, sh, opart, part=api.partition(r, 3, 32, 16))
        assert np.array_equal(acc, opart), r          # pixels of other ranks untouched, own pixels equal
    assert np.array_equal(acc, whole)


def test_arbitrary_rays(ctx, oracle, dsphere, sphere100):
    rng = np.random.default_rng(5)
    eyes = rng.uniform(-150, 150, (4000, 3))
    dirs = rng.normal(size=(4000, 3))
    dirs[::7, 1] = 0.0
    dirs[::11, 2] = -0.0
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    rays = refapi.make_rays(eyes, dirs)
    for space in (abi.SPACE_INDEX, abi.SPACE_WORLD):
        got = refapi.hits_to_dict(ctx.intersect(dsphere, rays, space=space), len(rays))
        want = oracle.intersect(sphere100.oracle_handle, rays, space=space)
        assert want["hit"].sum() > 500
        assert np.array_equal(got["hit"], want["hit"]) and np.array_equal(got["ijk"], want["ijk"])
        assert got.tobytes() == want.tobytes()
    # empty batch
    assert len(ctx.intersect(dsphere, refapi.rays_array(0))) == 0


def test_reference_kats_on_gpu(ctx, ref):
    """TestLevelSetRayIntersector.cc:37-120 through the GPU: sphere r=5 at (20,0,0), dx=0.5 -> xyz (15,0,0), t = 13"""
    g = ref.sphere(5.0, (20.0, 0.0, 0.0), 0.5, 2.0)
    dg = ctx.upload(ref.nanovdb(g))
    rays = refapi.make_rays([(2.0, 0.0, 0.0), (2.0, 0.0, 0.0), (2.0, 0.0, 0.0)], [(1.0, 0.0, 0.0), (1.0, -0.0, -0.0), (-1.0, 0.0, 0.0)])
    h = refapi.hits_to_dict(ctx.intersect(dg, rays), 3)
    assert list(h["hit"]) == [1, 1, 0]
    for k in (0, 1):
        assert np.allclose(h["xyz_world"][k], (15.0, 0.0, 0.0), atol=1e-6) and abs(h["t_world"][k] - 13.0) < 1e-6
    assert h[2].tobytes() == bytes(h[2].nbytes)      # a miss leaves the record untouched (zero)
    assert h.tobytes() == ref.intersect(g, rays).tobytes()
    dg.free()


def test_volume_vs_oracle_and_reference(ctx, ref, oracle, dfog, fog100):
    W, H = 160, 120
    d = refapi.camera_desc(W, H, translation=(60, 40, 280), lookat=(0, 0, 0))
    cam = api.vdb_render_camera(W, H, (60, 40, 280), (0, 0, 0))
    for pstep, sstep in ((0.5, 3.0), (1.0, 3.0), (2.0, 1.5)):
        vo = api.vol_opts_default()
        vo.primary_step, vo.shadow_step = pstep, sstep
        film = refapi.new_film(W, H)
        ctx.render_volume(dfog, cam, vo, film)
        ofilm = refapi.new_film(W, H)
        oracle.render_volume(fog100.oracle_handle, cam, vo, ofilm, threads=4)
        rfilm = refapi.new_film(W, H)
        ref.render_volume(fog100.ref_handle, d, vo, rfilm)
        assert np.array_equal(ofilm, rfilm)
        assert (film[..., 3] > 0).sum() > 3000
        assert np.array_equal(film[..., 3] > 0, ofilm[..., 3] > 0), "fog hit mask (alpha > 0)"
        assert np.allclose(film, ofilm, rtol=RTOL, atol=ATOL)
        mism = float((film != ofilm).any(axis=2).mean())
        print("fog step %.1f/%.1f: %.4f%% pixels not bit-identical" % (pstep, sstep, 100 * mism))


def test_volume_spans(ctx, oracle, dfog, fog100):
    W = H = 96
    cam = api.vdb_render_camera(W, H, (0, 0, 300), (0, 0, 0))
    rays = oracle.camera_rays(cam, [(i, j) for j in range(H) for i in range(0, W, 2)])
    s1, c1 = oracle.volume_spans(fog100.oracle_handle, rays)
    s2, c2 = ctx.volume_spans(dfog, rays)
    assert c1.max() >= 1 and np.array_equal(c1, c2) and np.array_equal(s1, s2)


def test_golden_fixtures(ctx):
    """outputs of the unmodified reference committed under tests/golden (make_golden.py), reproduced from a grid the
    library builds itself on the GPU"""
    gold = np.load(os.path.join(GOLD, "ls_sphere40.npz"))
    g = ctx.build_sphere(40.0, (3.0, -2.0, 1.0))
    W, H = 96, 72
    cam = api.vdb_render_camera(W, H, (30.0, 20.0, 140.0), (0, 0, 0))
    for name, kind in (("diffuse", abi.SHADER_DIFFUSE), ("normal", abi.SHADER_NORMAL), ("matte", abi.SHADER_MATTE)):
        film, aux = gpu_levelset(ctx, g, cam, api.make_shader(kind, (0.9, 0.8, 0.7, 1.0)), W, H, (0.1, 0.2, 0.3, 0.5))
        assert np.array_equal(film, gold["film_" + name]), name
    for k in ("hit", "ijk", "t_index", "t_world", "xyz", "nml"):
        assert np.array_equal(getattr(aux, k), gold[k]), k
    film = refapi.new_film(W, H)
    ctx.render_levelset(g, cam, api.make_shader(), film, spp=4, seed=0)
    assert np.array_equal(film, gold["film_spp4"])
    fgold = np.load(os.path.join(GOLD, "fog_sphere40.npz"))
    fog = ctx.build_fog(g)
    W, H = 64, 48
    cam = api.vdb_render_camera(W, H, (30.0, 20.0, 140.0), (0, 0, 0))
    vo = api.vol_opts_default()
    vo.primary_step = 0.5
    film = refapi.new_film(W, H)
    ctx.render_volume(fog, cam, vo, film)
    assert np.array_equal(film[..., 3] > 0, fgold["film"][..., 3] > 0)
    assert np.allclose(film, fgold["film"], rtol=RTOL, atol=ATOL)
    oracle = refapi.Oracle()
    rays = oracle.camera_rays(cam, [(i, j) for j in range(0, H, 4) for i in range(0, W, 4)])
    spans, counts = ctx.volume_spans(fog, rays)
    assert np.array_equal(counts, fgold["counts"]) and np.array_equal(spans, fgold["spans"])
    fog.free(); g.free()


def test_error_codes(ctx, dsphere, dfog):
    W, H = 16, 16
    cam = api.vdb_render_camera(W, H, (0, 0, 300), (0, 0, 0))
    film = refapi.new_film(W, H)
    sh = api.make_shader()
    with pytest.raises(api.VdbrtError) as e:
        ctx.render_levelset(dsphere, cam, sh, film, iso=3.0)
    assert e.value.code == 7
    with pytest.raises(api.VdbrtError) as e:
        ctx.render_levelset(dfog, cam, sh, film)
    assert e.value.code == 7          # background 0: the LinearSearchImpl member throws first, like the reference
    with pytest.raises(api.VdbrtError) as e:
        ctx.render_levelset(dsphere, cam, sh, film, spp=0)
    assert e.value.code == 8
    with pytest.raises(api.VdbrtError) as e:
        ctx.upload(np.zeros(4096, np.uint8))
    assert e.value.code == 2
    bad = dsphere.download().copy()
    bad[636] = 2                      # GridType::Double
    with pytest.raises(api.VdbrtError) as e:
        ctx.upload(bad)
    assert e.value.code == 3


def test_work_counters_match_oracle(ctx, oracle, dsphere, sphere100):
    W, H = 256, 256
    cam = api.vdb_render_camera(W, H, (0, 0, 300), (0, 0, 0))
    c = ctx.count_levelset(dsphere, cam).as_dict()
    film = refapi.new_film(W, H)
    _, eager = oracle.render_levelset(sphere100.oracle_handle, cam, api.make_shader(), film, counters=True)
    # the kernels evaluate tester.init's value on demand (lsAdvance); the oracle counts that way with its lazy-init knob, which
    # tests/test_oracle_vs_reference.py::test_lazy_tester_init_is_exact pins against the reference
    try:
        oracle.set_lazy_init(True)
        _, oc = oracle.render_levelset(sphere100.oracle_handle, cam, api.make_shader(), film, counters=True)
    finally:
        oracle.set_lazy_init(False)
    o = oc.as_dict()
    for k in ("rays", "root_probes", "upper_probes", "lower_probes", "voxel_probes", "hits"):
        assert c[k] == o[k] == eager.as_dict()[k], k
    # the GPU starts every ray with a cold stencil, the CPU keeps it across pixels: refills differ by < 1 %
    assert abs(c["stencil_refills"] - o["stencil_refills"]) < 0.01 * o["stencil_refills"]
    assert o["stencil_refills"] < eager.as_dict()["stencil_refills"]


def test_config2_full_size_vs_oracle(ctx, oracle):
    """BASELINE config 2 at full size: GPU-built torus R=650 r=325 (50 M active voxels, 0.58 GB), 1920x1080, Diffuse and Normal.
    Every pixel's hit flag, first-hit voxel, time, position, normal and colour must equal the oracle's bit for bit."""
    g = ctx.build_torus(650.0, 325.0)
    assert g.info.active_voxels == 50038096 and g.info.leaf_count == 257488
    og = oracle.open(g.download())
    W, H = 1920, 1080
    cam = api.vdb_render_camera(W, H, (0.0, 1.5 * 650, 3.0 * (650 + 325)), (0, 0, 0))
    threads = os.cpu_count() or 4
    for kind in (abi.SHADER_DIFFUSE, abi.SHADER_NORMAL):
        film, aux = gpu_levelset(ctx, g, cam, api.make_shader(kind), W, H)
        ofilm = refapi.new_film(W, H)
        oaux, _ = oracle.render_levelset(og, cam, api.make_shader(kind), ofilm, aux=True, threads=threads)
        assert int(aux.hit.sum()) == 1064534
        assert_records_equal(aux, oaux)
        assert np.array_equal(film, ofilm)
    g.free()


def test_config3_fog_vs_oracle(ctx, oracle):
    """BASELINE config 3 grid (fog sphere r=509 in a 1024^3 box, GPU-built, 0.19 GB) at a quarter of the resolution per axis:
    alpha>0 mask exact, colours within tolerance (measured: bit-identical)"""
    ls = ctx.build_sphere(509.0)
    fog = ctx.build_fog(ls)
    ls.free()
    og = oracle.open(fog.download())
    W, H = 480, 270
    cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3 * 509.0), (0, 0, 0))
    vo = api.vol_opts_default()
    vo.primary_step = 0.5
    film = refapi.new_film(W, H)
    ctx.render_volume(fog, cam, vo, film)
    ofilm = refapi.new_film(W, H)
    oracle.render_volume(og, cam, vo, ofilm, threads=os.cpu_count() or 4)
    assert (film[..., 3] > 0).sum() > 50000
    assert np.array_equal(film[..., 3] > 0, ofilm[..., 3] > 0)
    assert np.allclose(film, ofilm, rtol=RTOL, atol=ATOL)
    print("C3 fog: %.5f%% pixels not bit-identical" % (100.0 * float((film != ofilm).any(axis=2).mean())))
    fog.free()


def test_many_root_tiles_fall_back_to_global_table(ctx, oracle):
    """more root tiles than fit the shared-memory staging area (96): the kernels search the table in global memory instead"""
    rng = np.random.default_rng(7)
    n = 260
    s = np.column_stack([rng.uniform(-9000, 9000, (n, 3)), rng.uniform(12, 30, n)])
    s[0] = (0.0, 0.0, 0.0, 40.0)
    g = ctx.build_spheres(s)
    assert g.info.root_tiles > 96
    og = oracle.open(g.download())
    W, H = 160, 120
    cam = api.vdb_render_camera(W, H, (3000.0, 2000.0, 30000.0), (0, 0, 0))
    film, aux = gpu_levelset(ctx, g, cam, api.make_shader(), W, H)
    ofilm = refapi.new_film(W, H)
    oaux, _ = oracle.render_levelset(og, cam, api.make_shader(), ofilm, aux=True, threads=4)
    assert_records_equal(aux, oaux)
    assert np.array_equal(film, ofilm)
    # and a close-up of the sphere at the origin so that something is hit
    cam = api.vdb_render_camera(W, H, (20.0, 30.0, 150.0), (0, 0, 0))
    film, aux = gpu_levelset(ctx, g, cam, api.make_shader(), W, H)
    ofilm = refapi.new_film(W, H)
    oaux, _ = oracle.render_levelset(og, cam, api.make_shader(), ofilm, aux=True, threads=4)
    assert aux.hit.sum() > 2000
    assert_records_equal(aux, oaux)
    assert np.array_equal(film, ofilm)
    g.free()


def test_translated_grid_and_camera_inside(ctx, ref, oracle, sphere100):
    """a ScaleTranslateMap (index->world translation patched into the NanoVDB map) and a camera inside the narrow band's bbox"""
    buf = sphere100.buf.copy()
    t = np.array([10.5, -3.25, 7.0])
    buf[528:552] = np.frombuffer(t.astype("<f8").tobytes(), np.uint8)           # Map::mVecD
    buf[296 + 72:296 + 84] = np.frombuffer(t.astype("<f4").tobytes(), np.uint8)  # Map::mVecF
    buf = refapi.aligned_copy(buf)
    rg = ref.from_nanovdb(buf)
    og = oracle.open(buf)
    g = ctx.upload(buf)
    assert list(g.info.translation) == list(t)
    W, H = 200, 150
    for tr, look in (((30.0, 40.0, 290.0), tuple(t)), ((10.5, -3.25, 7.0), (100.0, 20.0, 30.0)), ((60.0, 0.0, 80.0), (200.0, 0.0, 100.0))):
        d = refapi.camera_desc(W, H, translation=tr, lookat=look)
        cam = api.vdb_render_camera(W, H, tr, look)
        film, aux = gpu_levelset(ctx, g, cam, api.make_shader(abi.SHADER_POSITION, bbox_min=(-100, -100, -100), inv_dim=(0.005,) * 3), W, H)
        ofilm = refapi.new_film(W, H)
        oaux, _ = oracle.render_levelset(og, cam, api.make_shader(abi.SHADER_POSITION, bbox_min=(-100, -100, -100), inv_dim=(0.005,) * 3), ofilm, aux=True)
        rfilm = refapi.new_film(W, H)
        ref.render_levelset(rg, d, refapi.shader(abi.SHADER_POSITION, bbox_min=(-100, -100, -100), inv_dim=(0.005,) * 3), rfilm)
        assert aux.hit.sum() > 1000
        assert_records_equal(aux, oaux)
        assert np.array_equal(film, ofilm)
        assert np.array_equal(film, rfilm)
    g.free()
