"""Colour-grid shaders: MatteShader / NormalShader / PositionShader / DiffuseShader with GridT = Vec3SGrid and
the default PointSampler (tools/RayTracer.h:542-725).  CPU part: the oracle port against the unmodified reference and against
a committed golden frame; GPU part: the CUDA path against the oracle and the golden frame."""
import os

import numpy as np
import pytest

from openvdb_b200 import api, _abi as abi
from tests import refapi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "color_shaders.npz")
W, H = 160, 120
KINDS = (("matte", abi.SHADER_MATTE), ("normal", abi.SHADER_NORMAL), ("position", abi.SHADER_POSITION), ("diffuse", abi.SHADER_DIFFUSE))


def shader(kind):
    return api.make_shader(kind, (0.3, 0.6, 0.9, 0.5), bbox_min=(-25.0, -28.0, -22.0), inv_dim=(1 / 50.0, 1 / 56.0, 1 / 44.0))


def camera():
    return api.vdb_render_camera(W, H, (18.0, 25.0, 90.0), (1.0, 2.0, 3.0))


def test_oracle_matches_reference(ref, oracle):
    """colour grid with its own transform (voxel 2, translated), voxels + one tile + background"""
    ls = ref.sphere(20.0, (1.0, 2.0, 3.0))
    col = ref.color_grid(ls, 2.0, (0.5, -0.25, 0.0))
    og, oc = oracle.open(ref.nanovdb(ls)), oracle.open_color(ref.color_nanovdb(col))
    d = refapi.camera_desc(W, H, translation=(18.0, 25.0, 90.0), lookat=(1.0, 2.0, 3.0))
    bg = (0.05, 0.1, 0.15, 0.7)
    for name, kind in KINDS:
        for spp in (1, 3):
            want = refapi.new_film(W, H, bg)
            ref.render_levelset_color(ls, col, d, shader(kind), want, spp=spp, seed=2, threaded=False)
            got = refapi.new_film(W, H, bg)
            oracle.render_levelset(og, camera(), shader(kind), got, spp=spp, jitter=api.jitter_table(2), color=oc)
            assert (want[..., :3] != np.float32(bg[:3])).any(axis=2).sum() > 2000
            assert np.array_equal(got, want), (name, spp)
    # the colour grid really matters: a constant-colour render differs
    plain = refapi.new_film(W, H, bg)
    oracle.render_levelset(og, camera(), shader(abi.SHADER_DIFFUSE), plain)
    assert not np.array_equal(plain, got)


def test_oracle_matches_golden(oracle):
    z = np.load(GOLD)
    ls, col = refapi.aligned_copy(z["ls"]), refapi.aligned_copy(z["color"])
    og, oc = oracle.open(ls), oracle.open_color(col)
    for name, kind in KINDS:
        got = refapi.new_film(W, H)
        oracle.render_levelset(og, camera(), shader(kind), got, color=oc)
        assert np.array_equal(got, z["film_" + name]), name


@pytest.mark.gpu
def test_gpu_matches_oracle_and_golden(ctx, oracle):
    z = np.load(GOLD)
    ls, col = refapi.aligned_copy(z["ls"]), refapi.aligned_copy(z["color"])
    og, oc = oracle.open(ls), oracle.open_color(col)
    g, c = ctx.upload(ls), ctx.upload_color(col)
    bg = (0.05, 0.1, 0.15, 0.7)
    for name, kind in KINDS:
        sh = api.make_shader(kind, (0.3, 0.6, 0.9, 0.5), bbox_min=(-25.0, -28.0, -22.0), inv_dim=(1 / 50.0, 1 / 56.0, 1 / 44.0), color_grid=c)
        film = refapi.new_film(W, H)
        ctx.render_levelset(g, camera(), sh, film)
        assert np.array_equal(film, z["film_" + name]), name
        for spp, rounds in ((1, True), (4, None)):
            got = refapi.new_film(W, H, bg)
            ctx.render_levelset(g, camera(), sh, got, opts=ctx.ls_opts(spp=spp, seed=2, rounds=rounds))
            want = refapi.new_film(W, H, bg)
            oracle.render_levelset(og, camera(), shader(kind), want, spp=spp, jitter=api.jitter_table(2), color=oc)
            assert np.array_equal(got, want), (name, spp)
    # misuse is refused
    with pytest.raises(api.VdbrtError):
        ctx.render_levelset(c, camera(), shader(abi.SHADER_DIFFUSE), refapi.new_film(W, H))      # a colour grid is not renderable
    with pytest.raises(api.VdbrtError):
        ctx.render_levelset(g, camera(), api.make_shader(abi.SHADER_DIFFUSE, color_grid=g), refapi.new_film(W, H))   # not a Vec3f grid
    with pytest.raises(api.VdbrtError):
        ctx.upload_color(ls)
    g.free(); c.free()
