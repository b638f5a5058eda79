"""The library's GPU grid builders against the reference's generators: same tree topology, same voxel values and active
states (checked through nanovdb::tools::nanoToOpenVDB + random/near-band probes), byte-size-identical buffers."""
import numpy as np
import pytest

from openvdb_b200 import api, _abi as abi
from tests import refapi

pytestmark = pytest.mark.gpu


def compare(ref, oracle, mine, refh, nprobe=200000, seed=0):
    buf = mine.download()
    st = ref.stats(refh)
    rb = ref.nanovdb(refh)
    assert buf.size == rb.size
    oi = oracle.info(oracle.open(rb))
    i = mine.info
    assert (i.leaf_count, i.lower_count, i.upper_count, i.root_tiles) == (oi.leaf_count, oi.lower_count, oi.upper_count, oi.root_tiles)
    assert list(i.node_bbox) == list(st["node_bbox"])
    mh = ref.from_nanovdb(buf)
    st2 = ref.stats(mh)
    assert st2["active_voxels"] == st["active_voxels"] and st2["leaf_count"] == st["leaf_count"] and st2["active_tiles"] == st["active_tiles"]
    assert st2["background"] == st["background"]
    rng = np.random.default_rng(seed)
    nb = np.array(st["node_bbox"])
    ijk = rng.integers(nb[:3] - 20, nb[3:] + 20, size=(nprobe, 3)).astype(np.int32)
    v1, a1 = ref.probe(refh, ijk)
    v2, a2 = ref.probe(mh, ijk)
    assert np.array_equal(v1, v2) and np.array_equal(a1, a2)
    act = ijk[a1 > 0][:40000]
    nbr = (act[:, None, :] + rng.integers(-2, 3, size=(len(act), 4, 3))).reshape(-1, 3).astype(np.int32)
    v1, a1 = ref.probe(refh, nbr)
    v2, a2 = ref.probe(mh, nbr)
    assert np.array_equal(v1, v2) and np.array_equal(a1, a2)
    # the oracle port reads the GPU-built buffer exactly like the reference-built one
    o1, b1 = oracle.probe(oracle.open(buf), nbr)
    assert np.array_equal(o1, v1) and np.array_equal(b1, a1)
    ref.free(mh)


def test_sphere(ctx, ref, oracle):
    g = ctx.build_sphere(100.0)
    assert g.info.leaf_count == 4025 and g.info.active_voxels == 753990 and g.info.bytes == 11064704
    compare(ref, oracle, g, ref.sphere(100.0))
    g.free()
    g = ctx.build_sphere(5.0, (20, 0, 0), 0.5, 2.0)
    compare(ref, oracle, g, ref.sphere(5.0, (20, 0, 0), 0.5, 2.0), 50000)
    g.free()
    g = ctx.build_sphere(33.3, (-71.5, 12.25, 40.0), 1.0, 3.0)
    compare(ref, oracle, g, ref.sphere(33.3, (-71.5, 12.25, 40.0), 1.0, 3.0), 100000)
    g.free()


def test_torus(ctx, ref, oracle):
    g = ctx.build_torus(100.0, 50.0)
    assert g.info.active_voxels == 1183940 and g.info.leaf_count == 6337
    compare(ref, oracle, g, ref.torus(100.0, 50.0))
    g.free()


def test_fog(ctx, ref, oracle):
    ls = ctx.build_sphere(100.0)
    fog = ctx.build_fog(ls)
    rls = ref.sphere(100.0)
    compare(ref, oracle, fog, ref.fog_from_levelset(rls))
    assert fog.info.grid_class == abi.GRID_CLASS_FOG_VOLUME and fog.info.background == 0.0
    fog.free(); ls.free()


def test_sphere_union(ctx, ref, oracle):
    rng = np.random.default_rng(20240607)
    s = np.column_stack([rng.uniform(-150, 150, (24, 3)), rng.uniform(10, 40, 24)])
    g = ctx.build_spheres(s)
    compare(ref, oracle, g, ref.spheres_union(s))
    g.free()


def test_built_grid_renders_like_reference_grid(ctx, ref):
    """end to end: GPU-built torus rendered by the GPU == reference-built torus rendered by the reference"""
    W, H = 240, 135
    g = ctx.build_torus(60.0, 25.0)
    cam = api.vdb_render_camera(W, H, (0, 90, 255), (0, 0, 0))
    film = refapi.new_film(W, H)
    ctx.render_levelset(g, cam, api.make_shader(abi.SHADER_NORMAL), film)
    rfilm = refapi.new_film(W, H)
    ref.render_levelset(ref.torus(60.0, 25.0), refapi.camera_desc(W, H, translation=(0, 90, 255), lookat=(0, 0, 0)),
                        refapi.shader(abi.SHADER_NORMAL), rfilm)
    assert (film[..., :3].sum(axis=2) > 0).sum() > 3000
    assert np.array_equal(film, rfilm)
    g.free()
