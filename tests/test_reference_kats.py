"""Replays the reference's own known-answer tests through the oracle port:
TestRay.cc (bbox clip, DDA stepping), TestLevelSetRayIntersector.cc (analytic sphere hits, misses leave outputs alone),
TestVolumeRayIntersector.cc (leaf-granular spans).  Numbers are transcribed from the cited lines of
openvdb/openvdb/unittest/*.cc; grids are made with the reference's generators (oracle/_ref)."""
import numpy as np
import pytest

from openvdb_b200 import api, _abi as abi
from tests import refapi

DELTA = 1e-9          # math::Delta<double>::value()
DBL_MAX = float(np.finfo(np.float64).max)


def ray(eye, d, t0=DELTA, t1=DBL_MAX):
    return refapi.make_rays([eye], [d], t0, t1)[0]


def test_ray_bbox_clip(oracle):
    """TestRay.cc:231-257: eye(2,1,1) dir(-1,2,3) normalised, CoordBBox((0,2,2),(2,4,6)) -> t0 = 0.5, t1 = 1.5 in units of
    |dir| (max NOT padded: the reference uses CoordBBox::max() as is)"""
    d = np.array([-1.0, 2.0, 3.0])
    n = np.linalg.norm(d)
    r = ray((2, 1, 1), d / n)
    hit, t0, t1 = oracle.ray_clip(r, [0, 2, 2, 2, 4, 6])
    assert hit
    assert abs(t0 - 0.5 * n) < 1e-12 and abs(t1 - 1.5 * n) < 1e-12
    # a miss leaves the times untouched (TestRay.cc:183-206)
    hit, _, _ = oracle.ray_clip(ray((2, 1, 1), (1.0, 0.0, 0.0)), [0, 2, 2, 2, 4, 6])
    assert not hit


def test_dda_first_steps(oracle):
    """TestRay.cc:313-346: DDA<Ray,12> from the origin along +x: time()==Delta, next()==1... wait for the big node:
    voxel (0,0,0), first next() is the exit of the 4096^3 node"""
    tr = oracle.dda_trace(ray((0, 0, 0), (1.0, 0.0, 0.0)), 12, 3)
    assert tr[0][0] == DELTA and tr[0][1] == 4096.0 and list(tr[0][2:]) == [0, 0, 0]
    assert tr[1][0] == 4096.0 and tr[1][1] == 8192.0 and list(tr[1][2:]) == [4096, 0, 0]


@pytest.mark.parametrize("axis,sign", [(a, s) for a in range(3) for s in (1.0, -1.0)])
def test_dda_unit_stepping(oracle, axis, sign):
    """TestRay.cc:348-452: unit-speed stepping at Log2Dim 0 and 3 along all six axis directions, +-0 in the other components"""
    for zero in (0.0, -0.0):
        d = [zero, zero, zero]
        d[axis] = sign
        eye = [0.5, 0.5, 0.5]
        tr = oracle.dda_trace(ray(eye, d), 0, 8)
        for k in range(1, len(tr)):
            assert abs(tr[k][0] - (k - 0.5)) < 1e-12                 # crosses a voxel face every unit of time
            assert tr[k][2 + axis] == (k if sign > 0 else -k)
            assert all(tr[k][2 + o] == 0 for o in range(3) if o != axis)
        tr = oracle.dda_trace(ray(eye, d), 3, 4)
        step = 8 if sign > 0 else -8
        assert tr[1][2 + axis] == step and abs(tr[1][0] - (7.5 if sign > 0 else 0.5)) < 1e-12
        assert abs(tr[2][0] - tr[1][0] - 8.0) < 1e-12


def sphere_case(ref, oracle, radius, center, dx, hw):
    g = ref.sphere(radius, center, dx, hw)
    buf = ref.nanovdb(g)
    return g, oracle.open(buf), buf


def test_levelset_intersector_analytic(ref, oracle):
    """TestLevelSetRayIntersector.cc:37-120: sphere r=5 at (20,0,0), dx=0.5, half-width 2; ray from (2,0,0) along +x hits at
    xyz=(15,0,0), t=13 (tolerance 1e-6 in the reference); also with -0.0 direction components and a start inside (t0=16)
    hitting the back face at x=25, t=21... (world-space times)"""
    g, og, buf = sphere_case(ref, oracle, 5.0, (20.0, 0.0, 0.0), 0.5, 2.0)
    for d in ((1.0, 0.0, 0.0), (1.0, -0.0, -0.0)):
        h = oracle.intersect(og, refapi.make_rays([(2.0, 0.0, 0.0)], [d]))[0]
        assert h["hit"] == 1
        assert np.allclose(h["xyz_world"], (15.0, 0.0, 0.0), atol=1e-6)
        assert abs(h["t_world"] - 13.0) < 1e-6
        assert np.allclose(h["nml"], (-1.0, 0.0, 0.0), atol=6e-2)     # box-stencil gradient at a voxel corner
    h = oracle.intersect(og, refapi.make_rays([(2.0, 0.0, 0.0)], [(1.0, 0.0, 0.0)], t0=16.0))[0]
    assert h["hit"] == 1 and np.allclose(h["xyz_world"], (25.0, 0.0, 0.0), atol=1e-6) and abs(h["t_world"] - 23.0) < 1e-6
    # the same through the reference itself (identical bits)
    r = ref.intersect(g, refapi.make_rays([(2.0, 0.0, 0.0)], [(1.0, 0.0, 0.0)]))[0]
    o = oracle.intersect(og, refapi.make_rays([(2.0, 0.0, 0.0)], [(1.0, 0.0, 0.0)]))[0]
    assert r.tobytes() == o.tobytes()


def test_levelset_intersector_other_voxel_sizes(ref, oracle):
    """TestLevelSetRayIntersector.cc:121-235: dx = 1.5 and a diagonal ray"""
    g, og, buf = sphere_case(ref, oracle, 5.0, (20.0, 0.0, 0.0), 1.5, 2.0)
    h = oracle.intersect(og, refapi.make_rays([(2.0, 0.0, 0.0)], [(1.0, 0.0, 0.0)]))[0]
    assert h["hit"] == 1 and np.allclose(h["xyz_world"], (15.0, 0.0, 0.0), atol=2e-2) and abs(h["t_world"] - 13.0) < 2e-2
    g, og, buf = sphere_case(ref, oracle, 5.0, (10.0, 10.0, 10.0), 0.5, 2.0)
    d = np.ones(3) / np.sqrt(3.0)
    h = oracle.intersect(og, refapi.make_rays([(0.0, 0.0, 0.0)], [d]))[0]
    t = np.sqrt(300.0) - 5.0
    assert h["hit"] == 1 and abs(h["t_world"] - t) < 1e-2 and np.allclose(h["xyz_world"], d * t, atol=1e-2)


def test_levelset_missed_intersections_leave_outputs(ref, oracle):
    """TestLevelSetRayIntersector.cc:311-389: on a miss nothing is written (here: the whole record stays zero / hit == 0)"""
    g, og, buf = sphere_case(ref, oracle, 5.0, (20.0, 0.0, 0.0), 0.5, 2.0)
    rays = refapi.make_rays([(2.0, 0.0, 0.0), (2.0, 30.0, 0.0)], [(-1.0, 0.0, 0.0), (1.0, 0.0, 0.0)])
    for space in (abi.SPACE_WORLD, abi.SPACE_INDEX):
        h = oracle.intersect(og, rays, space=space)
        assert not h["hit"].any()
        assert h.tobytes() == bytes(len(h.tobytes()))
        assert ref.intersect(g, rays, space=space).tobytes() == h.tobytes()


def test_levelset_sweep_accuracy(ref, oracle):
    """TestLevelSetRayIntersector.cc:236-309 (sweep over a dx=0.01-style sphere, scaled down): hit-time error < 0.1 %, position
    error < 0.06 voxel against the analytic sphere"""
    radius, dx = 2.0, 0.05
    g, og, buf = sphere_case(ref, oracle, radius, (0.0, 0.0, 0.0), dx, 3.0)
    n = 64
    u = (np.arange(n) + 0.5) / n * 2 - 1
    eyes = np.array([(x * 1.5, y * 1.5, 10.0) for y in u for x in u])
    dirs = np.tile((0.0, 0.0, -1.0), (len(eyes), 1))
    h = oracle.intersect(og, refapi.make_rays(eyes, dirs))
    r2 = eyes[:, 0] ** 2 + eyes[:, 1] ** 2
    inside = r2 < (radius - 2 * dx) ** 2
    assert h["hit"][inside].all()
    t_exact = 10.0 - np.sqrt(radius ** 2 - r2[inside])
    assert (np.abs(h["t_world"][inside] - t_exact) / t_exact).max() < 1e-3
    assert np.abs(np.linalg.norm(h["xyz_world"][inside], axis=1) - radius).max() < 0.06 * dx
    assert not h["hit"][r2 > (radius + 2 * dx) ** 2].any()


VOLUME_KATS = [
    # (voxels, boxes, eye, dir, expected spans)   TestVolumeRayIntersector.cc line refs in comments
    ([((0, 0, 0), 1.0), ((7, 7, 7), 1.0)], [], (-1, 0, 0), (1, 0, 0), [(1, 9)]),                                   # :37-52 single leaf
    ([((1, 1, 1), 1.0), ((7, 3, 3), 1.0)], [], (-1, 0, 0), (1, 0, 0), [(1, 9)]),                                   # :69-84
    ([((0, 0, 0), 1.0), ((8, 0, 0), 1.0), ((15, 7, 7), 1.0)], [], (-1, 0, 0), (1, 0, 0), [(1, 17)]),                # :101-117 adjacent leaves merge
    ([((0, 0, 0), 1.0), ((8, 0, 0), 1.0), ((24, 0, 0), 1.0), ((31, 7, 7), 1.0)], [], (-1, 0, 0), (1, 0, 0), [(1, 17), (25, 33)]),   # :118-140 gap
    ([((0, 0, 0), 1.0), ((8, 0, 0), 1.0), ((24, 0, 0), 1.0)], [((32, 0, 0), (39, 7, 7), 2.0, True)], (-1, 0, 0), (1, 0, 0), [(1, 17), (25, 41)]),  # :144-182 active tile appended
    ([((0, 0, 0), 1.0), ((8, 0, 0), 1.0), ((24, 0, 0), 1.0)], [], (50, 0, 0), (-1, 0, 0), [(18, 26), (34, 50)]),    # :202-220 "Jan": reversed ray
]


@pytest.mark.parametrize("case", range(len(VOLUME_KATS)))
def test_volume_intersector_spans(ref, oracle, case):
    voxels, boxes, eye, d, want = VOLUME_KATS[case]
    g = ref.custom(0.0, abi.GRID_CLASS_FOG_VOLUME, 1.0, voxels=voxels, boxes=boxes)
    og = oracle.open(ref.nanovdb(g))
    rays = refapi.make_rays([eye], [d])
    for impl in (lambda: ref.volume_spans(g, rays, space=abi.SPACE_INDEX), lambda: oracle.volume_spans(og, rays, space=abi.SPACE_INDEX)):
        spans, counts = impl()
        assert counts[0] == len(want)
        for k, (a, b) in enumerate(want):
            assert abs(spans[0, k, 0] - a) < 1e-6 and abs(spans[0, k, 1] - b) < 1e-6
    s1, c1 = ref.volume_spans(g, rays, space=abi.SPACE_INDEX)
    s2, c2 = oracle.volume_spans(og, rays, space=abi.SPACE_INDEX)
    assert np.array_equal(s1, s2) and np.array_equal(c1, c2)


def test_volume_bbox_hit_but_leaf_miss(ref, oracle):
    """TestVolumeRayIntersector.cc:221-250 ("Trevor"): the ray enters the node bbox but no leaf/tile is active on its path"""
    g = ref.custom(0.0, abi.GRID_CLASS_FOG_VOLUME, 1.0, voxels=[((0, 0, 0), 1.0), ((20, 20, 0), 1.0)])
    og = oracle.open(ref.nanovdb(g))
    rays = refapi.make_rays([(12.5, 4.5, 10.0)], [(0.0, 0.0, -1.0)])
    s1, c1 = ref.volume_spans(g, rays, space=abi.SPACE_INDEX)
    s2, c2 = oracle.volume_spans(og, rays, space=abi.SPACE_INDEX)
    assert c1[0] == 0 and c2[0] == 0


@pytest.mark.parametrize("iterations", [0, 1, 2, 3])
def test_levelset_search_iterations_match_the_reference(ref, oracle, iterations):
    """LinearSearchImpl<FloatGrid, Iterations> (tools/RayIntersector.h:630-636): the port's secant refinements are bit-identical to the
    stock LevelSetRayIntersector<FloatGrid, LinearSearchImpl<FloatGrid, N>> for world- and index-space rays, and so is a rendered frame"""
    radius, dx = 2.0, 0.05
    g, og, buf = sphere_case(ref, oracle, radius, (0.3, -0.2, 0.1), dx, 3.0)
    rng = np.random.default_rng(iterations)
    n = 4000
    eyes = np.column_stack([rng.uniform(-2.5, 2.5, n), rng.uniform(-2.5, 2.5, n), np.full(n, 10.0)])
    dirs = np.column_stack([rng.uniform(-0.05, 0.05, n), rng.uniform(-0.05, 0.05, n), np.full(n, -1.0)])
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    rays = refapi.make_rays(eyes, dirs)
    a = ref.intersect_iter(g, rays, iterations)
    b = oracle.intersect(og, rays, iterations=iterations)
    assert a["hit"].sum() > 1000
    for k in ("hit", "t_index", "t_world", "xyz_index", "xyz_world", "nml"):
        assert np.array_equal(a[k], b[k]), k
    if iterations:
        assert not np.array_equal(b["t_index"], oracle.intersect(og, rays)["t_index"])       # the refinement does change hit times
    # index-space rays (index outputs only on the reference side)
    irays = refapi.make_rays(eyes / dx, dirs)
    a = ref.intersect_iter(g, irays, iterations, space=abi.SPACE_INDEX)
    b = oracle.intersect(og, irays, space=abi.SPACE_INDEX, iterations=iterations)
    for k in ("hit", "t_index", "xyz_index"):
        assert np.array_equal(a[k], b[k]), k
    # a frame through tools::rayTrace(grid, intersector, ...)
    W, H = 96, 64
    d = refapi.camera_desc(W, H, translation=(1.0, 2.0, 9.0), lookat=(0, 0, 0))
    cam = api.vdb_render_camera(W, H, (1.0, 2.0, 9.0), (0, 0, 0))
    f_ref, f_port = refapi.new_film(W, H), refapi.new_film(W, H)
    ref.render_levelset_iter(g, d, refapi.shader(abi.SHADER_NORMAL), f_ref, iterations, spp=2, seed=4)
    oracle.render_levelset(og, cam, api.make_shader(abi.SHADER_NORMAL), f_port, spp=2, jitter=api.jitter_table(4), iterations=iterations)
    assert (f_ref[..., :3].sum(axis=2) > 0).sum() > 500
    assert np.array_equal(f_ref, f_port)
