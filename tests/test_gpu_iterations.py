"""LinearSearchImpl<GridT, Iterations> on the GPU (vdbrt_ls_opts::iterations, vdbrt_intersect_levelset_ex): bit-identical to the oracle
port (itself pinned to the unmodified reference for Iterations = 0..3, tests/test_reference_kats.py), and the reference's own accuracy
sweep with <FloatGrid, 2> (unittest/TestLevelSetRayIntersector.cc:279-309) replayed at its full size on the device."""
import numpy as np
import pytest

from openvdb_b200 import api, _abi as abi
from tests import refapi
from tests.test_gpu_parity import assert_records_equal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("iterations", [1, 2, 3, 7])
def test_iterations_match_oracle(ctx, oracle, torus_small, iterations):
    g = ctx.upload(torus_small.buf)
    rng = np.random.default_rng(iterations)
    n = 20000
    eyes = np.column_stack([rng.uniform(-90, 90, n), np.full(n, 120.0), rng.uniform(-90, 90, n)])
    dirs = np.column_stack([rng.uniform(-0.3, 0.3, n), np.full(n, -1.0), rng.uniform(-0.3, 0.3, n)])
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    rays = refapi.make_rays(eyes, dirs)
    for space in (abi.SPACE_WORLD, abi.SPACE_INDEX):
        got = refapi.hits_to_dict(ctx.intersect(g, rays, space=space, iterations=iterations), n)
        want = oracle.intersect(torus_small.oracle_handle, rays, space=space, iterations=iterations)
        assert want["hit"].sum() > 5000
        assert got.tobytes() == want.tobytes()
    assert not np.array_equal(want["t_index"], oracle.intersect(torus_small.oracle_handle, rays, space=abi.SPACE_INDEX)["t_index"])
    # frames: records and film, one and several samples per pixel, whole and partitioned
    W, H = 301, 187
    cam = api.vdb_render_camera(W, H, (0.0, 90.0, 255.0), (0, 0, 0))
    sh = api.make_shader(abi.SHADER_NORMAL, (0.9, 0.7, 0.5, 0.8))
    for spp in (1, 3):
        film, aux = refapi.new_film(W, H, (0.1, 0.2, 0.3, 1)), refapi.AuxArrays(W, H)
        ctx.render_levelset(g, cam, sh, film, aux=aux.pod(), opts=ctx.ls_opts(spp=spp, seed=2, iterations=iterations))
        ofilm = refapi.new_film(W, H, (0.1, 0.2, 0.3, 1))
        oaux, _ = oracle.render_levelset(torus_small.oracle_handle, cam, sh, ofilm, spp=spp, jitter=api.jitter_table(2), aux=True, threads=4, iterations=iterations)
        assert_records_equal(aux, oaux)
        assert np.array_equal(film, ofilm)
        film2 = refapi.new_film(W, H, (0.1, 0.2, 0.3, 1))
        for r in range(2):
            ctx.render_levelset(g, cam, sh, film2, opts=ctx.ls_opts(spp=spp, seed=2, iterations=iterations, part=api.partition(r, 2, 32, 16)))
        assert np.array_equal(film2, ofilm)
    g.free()


def test_reference_accuracy_sweep_with_two_iterations(ctx):
    """TestLevelSetRayIntersector.cc:279-309: sphere r = 5 at (10,10,20), voxel 0.01 (a 1000-voxel sphere), half-width 2,
    LinearSearchImpl<FloatGrid, 2>; 1024 x 1024 rays along +z: hit time within 0.1 % of the analytic one, hit position within 0.06 voxel"""
    r, c, s = 5.0, np.array([10.0, 10.0, 20.0]), 0.01
    g = ctx.build_sphere(r, tuple(c), voxel=s, half_width=2.0)
    width = 1024
    dx = 20.0 / width
    ii, jj = np.meshgrid(np.arange(width), np.arange(width), indexing="ij")
    eyes = np.column_stack([dx * ii.ravel(), dx * jj.ravel(), np.zeros(width * width)])
    dirs = np.tile((0.0, 0.0, 1.0), (width * width, 1))
    rays = refapi.make_rays(eyes, dirs)
    h = refapi.hits_to_dict(ctx.intersect(g, rays, iterations=2), width * width)
    hit = h["hit"] == 1
    assert hit.sum() > 190000
    d2 = (eyes[hit, 0] - c[0]) ** 2 + (eyes[hit, 1] - c[1]) ** 2
    assert (d2 <= r * r).all()                                 # EXPECT_TRUE(ray.intersects(c, r, t0, t1)) (tangent rays included)
    t0 = c[2] - np.sqrt(r * r - d2)
    assert np.abs(100 * (t0 - h["t_world"][hit]) / t0).max() < 0.1
    p0 = eyes[hit] + t0[:, None] * dirs[hit]
    assert (np.linalg.norm(p0 - h["xyz_world"][hit], axis=1) / s).max() < 0.06
    # and the refinement is what gets it there: without it the position error is larger
    h0 = refapi.hits_to_dict(ctx.intersect(g, rays), width * width)
    e2 = np.linalg.norm(p0 - h["xyz_world"][hit], axis=1).mean()
    e0 = np.linalg.norm(p0 - h0["xyz_world"][hit], axis=1).mean()
    assert e2 <= e0
    g.free()
