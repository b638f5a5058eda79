"""Long-ray rounds (csrc/vdbrt_kernels.cuh: suspend -> scout -> march -> resolve): the same pixels and per-pixel records as
the in-line traversal and as the oracle, whatever the budget and the number of rounds.

A leaf visit of the reference's LevelSetHDDA depends only on the ray and the visit's [t0,t1] (math/DDA.h:172-173), so
marching the leaves of one ray in parallel and taking the first hit in visit order must be bit-identical."""
import os

import numpy as np
import pytest

from openvdb_b200 import api, _abi as abi
from tests import refapi
from tests.test_gpu_parity import assert_records_equal

pytestmark = pytest.mark.gpu


def render(ctx, grid, cam, sh, W, H, bg=(0, 0, 0, 1), want_aux=True, **opts):
    film = refapi.new_film(W, H, bg)
    aux = refapi.AuxArrays(W, H)
    pod = aux.pod()
    ctx.render_levelset(grid, cam, sh, film, aux=pod if want_aux else None, opts=ctx.ls_opts(**opts))
    return film, aux


@pytest.fixture(scope="module")
def tiny_budget_ctx():
    """a context whose tiles may spend only a handful of iterations: nearly every ray that enters the grid is suspended"""
    old = {k: os.environ.get(k) for k in ("VDBRT_LS_BUDGET", "VDBRT_LS_FACTOR", "VDBRT_LS_LEAVES", "VDBRT_LS_TAIL")}
    os.environ.update(VDBRT_LS_BUDGET="6", VDBRT_LS_FACTOR="0", VDBRT_LS_LEAVES="2,4,12", VDBRT_LS_TAIL="0")    # the per-tile rule
    c = api.Context(0)
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    yield c
    c.close()


@pytest.mark.parametrize("kind", [abi.SHADER_DIFFUSE, abi.SHADER_NORMAL, abi.SHADER_POSITION])
def test_rounds_match_inline_and_oracle(ctx, oracle, torus_small, kind):
    """default budget: only the grazing rays of the torus silhouette go through the rounds"""
    g = ctx.upload(torus_small.buf)
    W, H = 400, 240
    cam = api.vdb_render_camera(W, H, (0.0, 90.0, 255.0), (0, 0, 0))
    sh = api.make_shader(kind, (0.9, 0.7, 0.5, 0.8), bbox_min=(-90, -30, -90), inv_dim=(1 / 180.0, 1 / 60.0, 1 / 180.0))
    bg = (0.1, 0.2, 0.3, 0.4)
    f_off, a_off = render(ctx, g, cam, sh, W, H, bg, rounds=False)
    f_on, a_on = render(ctx, g, cam, sh, W, H, bg, rounds=True)
    assert ctx.last_kernel_ms()[1] > 1            # the round kernels were launched
    ofilm = refapi.new_film(W, H, bg)
    oaux, _ = oracle.render_levelset(torus_small.oracle_handle, cam, sh, ofilm, aux=True, threads=4)
    assert a_on.hit.sum() > 10000
    assert_records_equal(a_on, a_off)
    assert_records_equal(a_on, oaux)
    assert np.array_equal(f_on, f_off) and np.array_equal(f_on, ofilm)
    f_noaux, _ = render(ctx, g, cam, sh, W, H, bg, want_aux=False, rounds=True)
    assert np.array_equal(f_noaux, ofilm)
    g.free()


def test_every_ray_suspended_few_rounds(tiny_budget_ctx, oracle, torus_small, sphere100):
    """budget 6, three rounds (K = 2, 4, 12): most rays are suspended at once, many outlive the rounds and are finished
    in line by k_long_finish; misses keep the (non-uniform) old film"""
    c = tiny_budget_ctx
    for gs, tr in ((torus_small, (0.0, 90.0, 255.0)), (sphere100, (30.0, 40.0, 290.0)), (torus_small, (0.0, 26.0, 200.0))):
        g = c.upload(gs.buf)
        W, H = 256, 160
        cam = api.vdb_render_camera(W, H, tr, (0, 0, 0))
        sh = api.make_shader(abi.SHADER_DIFFUSE)
        rng = np.random.default_rng(5)
        old = rng.random((H, W, 4)).astype(np.float32)
        film = old.copy()
        aux = refapi.AuxArrays(W, H)
        pod = aux.pod()
        c.render_levelset(g, cam, sh, film, aux=pod, opts=c.ls_opts(rounds=True))
        ofilm = old.copy()
        oaux, _ = oracle.render_levelset(gs.oracle_handle, cam, sh, ofilm, aux=True, threads=4)
        assert aux.hit.sum() > 3000
        assert_records_equal(aux, oaux)
        assert np.array_equal(film, ofilm)
        g.free()


def test_tail_rule_any_budget(ctx, oracle, torus_small):
    """the default rule: rays are suspended only once the queue has run dry, `ls_tail` iterations later -- from 'everything that
    is in flight at that moment' (1) to 'next to nothing' (4096), whole and partitioned frames give the oracle's frame and records"""
    g = ctx.upload(torus_small.buf)
    W, H = 400, 240
    cam = api.vdb_render_camera(W, H, (0.0, 90.0, 255.0), (0, 0, 0))
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    rng = np.random.default_rng(11)
    old = rng.random((H, W, 4)).astype(np.float32)
    ofilm = old.copy()
    oaux, _ = oracle.render_levelset(torus_small.oracle_handle, cam, sh, ofilm, aux=True, threads=4)
    try:
        for tail in (1, 7, 48, 4096):
            ctx.set_tuning(ls_tail=tail)
            film = old.copy()
            aux = refapi.AuxArrays(W, H)
            ctx.render_levelset(g, cam, sh, film, aux=aux.pod(), opts=ctx.ls_opts())
            assert ctx.last_kernel_ms()[1] > 1          # rounds are on by default
            assert_records_equal(aux, oaux)
            assert np.array_equal(film, ofilm), tail
            film = old.copy()
            for r in range(3):
                ctx.render_levelset(g, cam, sh, film, opts=ctx.ls_opts(part=api.partition(r, 3, 32, 32)))
            assert np.array_equal(film, ofilm), ("partitioned", tail)
    finally:
        ctx.set_tuning(ls_tail=24)
    g.free()


def test_partitioned_frame_uses_rounds_by_default(ctx, oracle, torus_small):
    """three ranks' tiles rendered one after the other into one device film (what bench.py does per rank)"""
    import torch
    g = ctx.upload(torus_small.buf)
    W, H = 384, 256
    cam = api.vdb_render_camera(W, H, (0.0, 90.0, 255.0), (0, 0, 0))
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    film = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    film[..., 3] = 1.0
    launches = []
    for r in range(3):
        o = ctx.ls_opts(part=api.partition(r, 3, 32, 32))
        ctx.render_levelset(g, cam, sh, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=o)
        launches.append(ctx.last_kernel_ms()[1])
    assert min(launches) > 1
    ofilm = refapi.new_film(W, H)
    oracle.render_levelset(torus_small.oracle_handle, cam, sh, ofilm, threads=4)
    assert np.array_equal(film.cpu().numpy(), ofilm)
    g.free()


def test_supersampling_never_uses_rounds(ctx, oracle, torus_small):
    """the samples of a pixel are accumulated in order by one thread (tools/RayTracer.h:903-915): rounds stay off"""
    g = ctx.upload(torus_small.buf)
    W, H = 160, 96
    cam = api.vdb_render_camera(W, H, (0.0, 90.0, 255.0), (0, 0, 0))
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    film = refapi.new_film(W, H)
    ctx.render_levelset(g, cam, sh, film, opts=ctx.ls_opts(spp=4, seed=1, rounds=True))
    assert ctx.last_kernel_ms()[1] == 1
    ofilm = refapi.new_film(W, H)
    oracle.render_levelset(torus_small.oracle_handle, cam, sh, ofilm, spp=4, jitter=api.jitter_table(1), threads=4)
    assert np.array_equal(film, ofilm)
    g.free()
