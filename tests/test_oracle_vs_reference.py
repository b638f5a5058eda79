"""Pins the oracle: the CPU restatement (oracle/vdbrt_oracle.cc) must agree BIT FOR BIT with the unmodified reference
(oracle/_ref/libvdbref.so = OpenVDB 13.0.1 LevelSetRayTracer / LevelSetRayIntersector / VolumeRender compiled from
/root/reference) on the same NanoVDB-serialised grids, and with the committed golden fixtures when the reference is
not available (GPU box)."""
import os

import numpy as np
import pytest

from openvdb_b200 import api, _abi as abi
from tests import refapi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def render_both(ref, oracle, gs, desc, sh, spp=1, seed=0, bg=(0, 0, 0, 1), iso=0.0):
    cam = ref.camera_pod(desc)
    f_ref = refapi.new_film(desc.width, desc.height, bg)
    ref.render_levelset(gs.ref_handle, desc, sh, f_ref, iso=iso, spp=spp, seed=seed, threaded=False)
    f_orc = refapi.new_film(desc.width, desc.height, bg)
    oracle.render_levelset(gs.oracle_handle, cam, sh, f_orc, iso=iso, spp=spp, jitter=ref.jitter_table(seed))
    return f_ref, f_orc


def test_grid_info_matches_reference(ref, oracle, sphere100, fog100):
    for gs in (sphere100, fog100):
        st = ref.stats(gs.ref_handle)
        info = oracle.info(gs.oracle_handle)
        assert list(info.node_bbox) == list(st["node_bbox"])          # leaf-granular bbox
        assert info.leaf_count == st["leaf_count"]
        assert info.background == st["background"]
    assert list(oracle.info(sphere100.oracle_handle).node_bbox) == [-104, -104, -104, 103, 103, 103]
    assert list(oracle.info(sphere100.oracle_handle).index_bbox) == [-102, -102, -102, 102, 102, 102]


@pytest.mark.parametrize("kind", [abi.SHADER_DIFFUSE, abi.SHADER_NORMAL, abi.SHADER_MATTE, abi.SHADER_POSITION])
def test_levelset_render_bit_exact(ref, oracle, sphere100, kind):
    d = refapi.camera_desc(192, 128, translation=(40, 60, 280), lookat=(0, 0, 0))
    sh = refapi.shader(kind, (0.9, 0.7, 0.5, 0.8), bbox_min=(-100, -100, -100), inv_dim=(1 / 200.0,) * 3)
    f_ref, f_orc = render_both(ref, oracle, sphere100, d, sh, bg=(0.1, 0.2, 0.3, 0.4))
    assert (f_ref[..., :3].sum(axis=2) != 0.6).sum() > 3000       # something was hit
    assert np.array_equal(f_ref, f_orc)


def test_levelset_config1_hit_count(ref, oracle, sphere100):
    """BASELINE config 1 at quarter resolution: sphere r=100, camera (0,0,300) -> origin, vdb_render lens"""
    d = refapi.camera_desc(256, 256, translation=(0, 0, 300), lookat=(0, 0, 0))
    f_ref, f_orc = render_both(ref, oracle, sphere100, d, refapi.shader())
    assert np.array_equal(f_ref, f_orc)
    assert int((f_ref[..., 0] > 0).sum()) == 37896               # 606 028 at 1024^2; 37 896 at 256^2


def test_levelset_records_bit_exact(ref, oracle, sphere100):
    d = refapi.camera_desc(160, 120, translation=(0, 0, 300), lookat=(0, 0, 0))
    raux, rctr, mism = ref.levelset_records(sphere100.ref_handle, d)
    assert mism == 0                                              # forwarding tester == stock intersector
    film = refapi.new_film(160, 120)
    oaux, octr = oracle.render_levelset(sphere100.oracle_handle, ref.camera_pod(d), refapi.shader(), film, aux=True, counters=True)
    for k in ("hit", "ijk", "t_index", "t_world", "xyz", "nml"):
        assert np.array_equal(getattr(raux, k), getattr(oaux, k)), k
    for k in ("root_probes", "upper_probes", "lower_probes", "voxel_probes", "stencil_refills"):
        assert getattr(rctr, k) == getattr(octr, k), k
    assert octr.hits == int(raux.hit.sum())


def test_lazy_tester_init_is_exact(ref, oracle, sphere100, torus_small, sphere_small, union_small):
    """The CUDA kernels evaluate tester.init's mV[0] only when the first voxel of a leaf visit passes the value gate (lsAdvance).  The
    same change made to the oracle must leave every record of every pixel as the REFERENCE has it, and only drop stencil refills."""
    d = refapi.camera_desc(160, 120, translation=(0, 0, 300), lookat=(0, 0, 0))
    raux, rctr, _ = ref.levelset_records(sphere100.ref_handle, d)
    try:
        oracle.set_lazy_init(True)
        film = refapi.new_film(160, 120)
        oaux, octr = oracle.render_levelset(sphere100.oracle_handle, ref.camera_pod(d), refapi.shader(), film, aux=True, counters=True)
        for k in ("hit", "ijk", "t_index", "t_world", "xyz", "nml"):
            assert np.array_equal(getattr(raux, k), getattr(oaux, k)), k
        for k in ("root_probes", "upper_probes", "lower_probes", "voxel_probes"):
            assert getattr(rctr, k) == getattr(octr, k), k
        assert octr.stencil_refills < rctr.stencil_refills
        # grazing rays (many leaf visits without a gated voxel), all shaders' inputs, two samples per pixel, refinements
        d2 = refapi.camera_desc(128, 96, translation=(0.0, 60.0, 120.0), lookat=(0, 0, 0))
        for iters in (0, 2):
            f_ref = refapi.new_film(128, 96)
            ref.render_levelset_iter(torus_small.ref_handle, d2, refapi.shader(abi.SHADER_NORMAL), f_ref, iters, spp=2, seed=3)
            f_orc = refapi.new_film(128, 96)
            oracle.render_levelset(torus_small.oracle_handle, ref.camera_pod(d2), refapi.shader(abi.SHADER_NORMAL), f_orc, spp=2,
                                   jitter=ref.jitter_table(3), iterations=iters)
            assert np.array_equal(f_ref, f_orc)
        # orthographic camera with iso != 0, a scaled + translated grid, a union of spheres: render_both compares reference and oracle films
        d3 = refapi.camera_desc(128, 96, translation=(10, 5, 250), rotation=(5, -10, 20), kind=abi.CAMERA_ORTHOGRAPHIC, frame=260.0)
        for iso in (1.25, -2.0):
            f_ref, f_orc = render_both(ref, oracle, sphere100, d3, refapi.shader(abi.SHADER_NORMAL), iso=iso)
            assert np.array_equal(f_ref, f_orc), iso
        d4 = refapi.camera_desc(128, 128, translation=(2, 3, 30), lookat=(20, 0, 0))
        f_ref, f_orc = render_both(ref, oracle, sphere_small, d4, refapi.shader())
        assert (f_ref[..., 0] > 0).sum() > 500 and np.array_equal(f_ref, f_orc)
        d5 = refapi.camera_desc(160, 120, translation=(30.0, 40.0, 150.0), lookat=(0, 0, 0))
        f_ref, f_orc = render_both(ref, oracle, union_small, d5, refapi.shader(abi.SHADER_POSITION, bbox_min=(-60, -60, -60), inv_dim=(1 / 120.0,) * 3))
        assert np.array_equal(f_ref, f_orc)
    finally:
        oracle.set_lazy_init(False)


def test_levelset_jittered_supersampling(ref, oracle, sphere100):
    d = refapi.camera_desc(96, 64, translation=(0, 0, 300), lookat=(0, 0, 0))
    for spp, seed in ((2, 0), (5, 3), (16, 0)):
        f_ref, f_orc = render_both(ref, oracle, sphere100, d, refapi.shader(), spp=spp, seed=seed)
        assert np.array_equal(f_ref, f_orc), (spp, seed)


def test_levelset_orthographic_and_iso(ref, oracle, sphere100):
    d = refapi.camera_desc(128, 96, translation=(10, 5, 250), rotation=(5, -10, 20), kind=abi.CAMERA_ORTHOGRAPHIC, frame=260.0)
    for iso in (0.0, 1.25, -2.0):
        f_ref, f_orc = render_both(ref, oracle, sphere100, d, refapi.shader(abi.SHADER_NORMAL), iso=iso)
        assert np.array_equal(f_ref, f_orc), iso


def test_levelset_scaled_translated_grid(ref, oracle, sphere_small):
    """dx = 0.5 sphere at (20,0,0): exercises worldToIndex / indexToWorld / applyIJT with a non-unit scale"""
    d = refapi.camera_desc(128, 128, translation=(2, 3, 30), lookat=(20, 0, 0))
    f_ref, f_orc = render_both(ref, oracle, sphere_small, d, refapi.shader())
    assert (f_ref[..., 0] > 0).sum() > 500
    assert np.array_equal(f_ref, f_orc)
    rays = ref.camera_rays(d, [(i, j) for j in range(0, 128, 3) for i in range(0, 128, 3)])
    a = ref.intersect(sphere_small.ref_handle, rays)
    b = oracle.intersect(sphere_small.oracle_handle, rays)
    assert a.tobytes() == b.tobytes()


def test_levelset_torus_and_union(ref, oracle, torus_small, union_small):
    d = refapi.camera_desc(160, 90, translation=(0, 90, 255), lookat=(0, 0, 0))
    f_ref, f_orc = render_both(ref, oracle, torus_small, d, refapi.shader())
    assert np.array_equal(f_ref, f_orc)
    d = refapi.camera_desc(160, 90, translation=(50, 120, 520), lookat=(0, 0, 0))
    f_ref, f_orc = render_both(ref, oracle, union_small, d, refapi.shader(abi.SHADER_NORMAL))
    assert (f_ref[..., :3].sum(axis=2) > 0).sum() > 1000
    assert np.array_equal(f_ref, f_orc)


def test_index_space_rays_and_inside_start(ref, oracle, sphere100):
    rng = np.random.default_rng(5)
    eyes = rng.uniform(-150, 150, (300, 3))
    dirs = rng.normal(size=(300, 3))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    dirs[::7, 1] = 0.0                    # exact zero components exercise the isZero branch of DDA::init
    dirs[::11, 2] = -0.0
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    rays = refapi.make_rays(eyes, dirs)
    for space in (abi.SPACE_INDEX, abi.SPACE_WORLD):
        a = ref.intersect(sphere100.ref_handle, rays, space=space)
        b = oracle.intersect(sphere100.oracle_handle, rays, space=space)
        assert a["hit"].sum() > 50
        assert a.tobytes() == b.tobytes()


def test_volume_render_bit_exact(ref, oracle, fog100):
    d = refapi.camera_desc(96, 72, translation=(60, 40, 280), lookat=(0, 0, 0))
    cam = ref.camera_pod(d)
    for pstep, sstep in ((0.5, 3.0), (1.0, 3.0), (2.0, 1.5)):
        vo = ref.vol_defaults()
        vo.primary_step, vo.shadow_step = pstep, sstep
        f_ref = refapi.new_film(96, 72)
        ref.render_volume(fog100.ref_handle, d, vo, f_ref)
        f_orc = refapi.new_film(96, 72)
        oracle.render_volume(fog100.oracle_handle, cam, vo, f_orc)
        assert (f_ref[..., 3] > 0).sum() > 1000
        assert np.array_equal(f_ref, f_orc), (pstep, sstep)


def test_volume_spans_bit_exact(ref, oracle, fog100):
    d = refapi.camera_desc(64, 64, translation=(0, 0, 300), lookat=(0, 0, 0))
    rays = ref.camera_rays(d, [(i, j) for j in range(64) for i in range(0, 64, 2)])
    s1, c1 = ref.volume_spans(fog100.ref_handle, rays)
    s2, c2 = oracle.volume_spans(fog100.oracle_handle, rays)
    assert c1.max() >= 1 and np.array_equal(c1, c2) and np.array_equal(s1, s2)


def test_error_conditions_match_reference(ref, oracle, sphere100, fog100):
    """the reference throws at construction; the restatement returns the matching code"""
    d = refapi.camera_desc(8, 8, translation=(0, 0, 300), lookat=(0, 0, 0))
    cam = ref.camera_pod(d)
    film = refapi.new_film(8, 8)
    # iso outside the narrow band -> ValueError (tools/RayIntersector.h:536-539)
    with pytest.raises(RuntimeError, match="ValueError"):
        ref.render_levelset(sphere100.ref_handle, d, refapi.shader(), film, iso=3.0)
    with pytest.raises(refapi.OracleError) as e:
        oracle.render_levelset(sphere100.oracle_handle, cam, refapi.shader(), film, iso=3.0)
    assert e.value.code == 7
    # a fog volume has background 0, so the LinearSearchImpl member throws first: ValueError (:536-539)
    with pytest.raises(RuntimeError, match="ValueError"):
        ref.render_levelset(fog100.ref_handle, d, refapi.shader(), film)
    with pytest.raises(refapi.OracleError) as e:
        oracle.render_levelset(fog100.oracle_handle, cam, refapi.shader(), film)
    assert e.value.code == 7
    # a grid that is not tagged as a level set (but has a usable background) -> RuntimeError (:105-109)
    g = ref.custom(3.0, abi.GRID_CLASS_FOG_VOLUME, 1.0, voxels=[((0, 0, 0), 1.0), ((1, 0, 0), -1.0)])
    og = oracle.open(ref.nanovdb(g))
    with pytest.raises(RuntimeError, match="RuntimeError"):
        ref.render_levelset(g, d, refapi.shader(), film)
    with pytest.raises(refapi.OracleError) as e:
        oracle.render_levelset(og, cam, refapi.shader(), film)
    assert e.value.code == 4
    # spp == 0 -> ValueError (tools/RayTracer.h:877-879)
    with pytest.raises(RuntimeError, match="ValueError"):
        ref.render_levelset(sphere100.ref_handle, d, refapi.shader(), film, spp=0)
    with pytest.raises(refapi.OracleError) as e:
        oracle.render_levelset(sphere100.oracle_handle, cam, refapi.shader(), film, spp=0)
    assert e.value.code == 8


def rotated_copy(buf, angle_deg=(20.0, -35.0, 50.0), scale=0.75, translation=(3.0, -2.0, 1.5)):
    """the same tree under a rotated, uniformly scaled, translated index->world map: nanovdb::Map (GridData + 296) holds float and double
    copies of the matrix, its inverse and the translation (NanoVDB.h:1418-1428); GridData::mVoxelSize is at +608"""
    ax, ay, az = np.radians(angle_deg)
    Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    A = scale * (Rz @ Ry @ Rx)                                   # world = A index + t
    Ai = np.linalg.inv(A)
    t = np.asarray(translation, np.float64)
    out = buf.copy()
    m = 296
    out[m:m + 36] = np.frombuffer(A.astype("<f4").tobytes(), np.uint8)
    out[m + 36:m + 72] = np.frombuffer(Ai.astype("<f4").tobytes(), np.uint8)
    out[m + 72:m + 84] = np.frombuffer(t.astype("<f4").tobytes(), np.uint8)
    out[m + 88:m + 160] = np.frombuffer(A.astype("<f8").tobytes(), np.uint8)
    out[m + 160:m + 232] = np.frombuffer(Ai.astype("<f8").tobytes(), np.uint8)
    out[m + 232:m + 256] = np.frombuffer(t.astype("<f8").tobytes(), np.uint8)
    out[608:632] = np.frombuffer(np.full(3, scale, "<f8").tobytes(), np.uint8)
    return refapi.aligned_copy(out), A, t


def test_rotated_affine_map_within_tolerance(ref, oracle, sphere100, fog100):
    """a grid whose index->world map has off-diagonal terms (OpenVDB AffineMap, math/Maps.h:411-445) is the TOLERANCE path:
    the port evaluates NanoVDB's stored matrices, the reference multiplies through its own 4x4s -- same frame within 1e-4 rel / 1e-3 abs,
    hit mask equal except for a handful of silhouette pixels"""
    buf, A, t = rotated_copy(sphere100.buf)
    rg = ref.from_nanovdb(buf)
    og = oracle.open(buf)
    assert abs(oracle.info(og).voxel_size[0] - 0.75) < 1e-12
    W, H = 240, 160
    centre = t
    eye = tuple(centre + np.array([30.0, 40.0, 250.0]))
    d = refapi.camera_desc(W, H, translation=eye, lookat=tuple(centre))
    cam = api.vdb_render_camera(W, H, eye, tuple(centre))
    for kind in (abi.SHADER_DIFFUSE, abi.SHADER_NORMAL):
        f_ref, f_port = refapi.new_film(W, H), refapi.new_film(W, H)
        ref.render_levelset(rg, d, refapi.shader(kind), f_ref)
        oracle.render_levelset(og, cam, api.make_shader(kind), f_port)
        hit_ref, hit_port = f_ref[..., :3].sum(axis=2) > 0, f_port[..., :3].sum(axis=2) > 0
        assert hit_ref.sum() > 5000
        assert (hit_ref != hit_port).sum() <= 4
        both = hit_ref & hit_port
        assert np.allclose(f_ref[both], f_port[both], rtol=1e-4, atol=1e-3)
    # arbitrary rays: times, positions and normals
    rng = np.random.default_rng(3)
    n = 3000
    eyes = centre + np.column_stack([rng.uniform(-60, 60, n), rng.uniform(-60, 60, n), np.full(n, 200.0)])
    dirs = np.column_stack([rng.uniform(-0.1, 0.1, n), rng.uniform(-0.1, 0.1, n), np.full(n, -1.0)])
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    rays = refapi.make_rays(eyes, dirs)
    a, b = ref.intersect(rg, rays), oracle.intersect(og, rays)
    assert a["hit"].sum() > 1000 and (a["hit"] != b["hit"]).sum() <= 2
    both = (a["hit"] == 1) & (b["hit"] == 1)
    for k in ("t_world", "xyz_world", "nml"):
        assert np.allclose(a[k][both], b[k][both], rtol=1e-4, atol=1e-3), k
    # fog through the same map
    fbuf, _, _ = rotated_copy(fog100.buf)
    rf, of = ref.from_nanovdb(fbuf), oracle.open(fbuf)
    vo = api.vol_opts_default()
    vo.primary_step = 0.5
    v_ref, v_port = refapi.new_film(W, H), refapi.new_film(W, H)
    ref.render_volume(rf, d, vo, v_ref)
    oracle.render_volume(of, cam, vo, v_port)
    assert (v_ref[..., 3] > 0).sum() > 5000
    assert ((v_ref[..., 3] > 0) != (v_port[..., 3] > 0)).sum() <= 4
    assert np.mean(np.abs(v_ref - v_port) > 1e-3 + 1e-4 * np.abs(v_ref)) < 2e-3
