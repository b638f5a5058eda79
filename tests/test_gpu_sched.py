"""How the render warps are fed (Sched, csrc/vdbrt_kernels.cuh) never changes a pixel: strips of tiles, lanes re-fed from the
warp's strip at any threshold, eager strip changes, heavy strips first (k_probe_levelset) -- every combination gives the
oracle's film and per-pixel records, for one and for several samples per pixel, whole and partitioned frames."""
import itertools

import numpy as np
import pytest

from openvdb_b200 import api, _abi as abi
from tests import refapi
from tests.test_gpu_parity import assert_records_equal

pytestmark = pytest.mark.gpu

DEFAULTS = dict(ls_strip=1, ls_strip_ratio=4, ls_refill=32, ls_eager=0, ls_affine=0, ls_order=0, ls_probe_cap=128, ls_probe_b=64, ls_history=1, ls_hist_a=250, ls_hist_b=105)


@pytest.fixture()
def tuned(ctx):
    ctx.set_tuning(ls_strip_ratio=0, ls_history=0)      # small test frames keep the strip length they ask for; history has its own test
    yield ctx
    ctx.set_tuning(**DEFAULTS)


def test_feeding_variants_match_oracle(tuned, oracle, torus_small):
    ctx = tuned
    g = ctx.upload(torus_small.buf)
    W, H = 403, 237                       # edge tiles with slots outside the film
    cam = api.vdb_render_camera(W, H, (0.0, 90.0, 255.0), (0, 0, 0))
    sh = api.make_shader(abi.SHADER_NORMAL, (0.9, 0.7, 0.5, 0.8))
    bg = (0.1, 0.2, 0.3, 0.4)
    ofilm = refapi.new_film(W, H, bg)
    oaux, _ = oracle.render_levelset(torus_small.oracle_handle, cam, sh, ofilm, aux=True, threads=4)
    assert oaux.hit.sum() > 10000
    for strip, refill, eager, order in itertools.product((1, 3, 8), (1, 8, 16, 32), (0, 1), (False, True)):
        ctx.set_tuning(ls_strip=strip, ls_refill=refill, ls_eager=eager, ls_probe_cap=24, ls_probe_b=8)
        for want_aux in (True, False):
            film = refapi.new_film(W, H, bg)
            aux = refapi.AuxArrays(W, H)
            ctx.render_levelset(g, cam, sh, film, aux=aux.pod() if want_aux else None, opts=ctx.ls_opts(order=order, rounds=False))
            assert ctx.last_kernel_ms()[1] == (2 if order else 1)   # rounds=False: the render kernel (+ the probe)
            assert np.array_equal(film, ofilm), (strip, refill, eager, order, want_aux)
            if want_aux:
                assert_records_equal(aux, oaux)
    g.free()


def test_feeding_variants_supersampled_and_partitioned(tuned, oracle, torus_small):
    ctx = tuned
    g = ctx.upload(torus_small.buf)
    W, H = 200, 120
    cam = api.vdb_render_camera(W, H, (0.0, 90.0, 255.0), (0, 0, 0))
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    ofilm = refapi.new_film(W, H)
    oracle.render_levelset(torus_small.oracle_handle, cam, sh, ofilm, spp=5, jitter=api.jitter_table(3), threads=4)
    for strip, refill, eager, order in ((8, 16, 0, True), (8, 4, 1, True), (2, 32, 0, True), (4, 12, 1, False)):
        ctx.set_tuning(ls_strip=strip, ls_refill=refill, ls_eager=eager, ls_probe_cap=16, ls_probe_b=6)
        film = refapi.new_film(W, H)
        ctx.render_levelset(g, cam, sh, film, opts=ctx.ls_opts(spp=5, seed=3, order=order))
        assert np.array_equal(film, ofilm), (strip, refill, eager, order)
        # three ranks, small macro tiles: every rank renders its share into the same film
        film = refapi.new_film(W, H)
        for r in range(3):
            ctx.render_levelset(g, cam, sh, film, opts=ctx.ls_opts(spp=5, seed=3, order=order, part=api.partition(r, 3, 16, 8)))
        assert np.array_equal(film, ofilm), ("partitioned", strip, refill, eager, order)
    g.free()


def test_ordering_puts_unfinished_probes_first(tuned, oracle, torus_small):
    """with a probe budget of a few steps nearly every strip that touches the grid is 'heavy': the frame is the same"""
    ctx = tuned
    g = ctx.upload(torus_small.buf)
    W, H = 320, 200
    cam = api.vdb_render_camera(W, H, (0.0, 90.0, 255.0), (0, 0, 0))
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    ofilm = refapi.new_film(W, H)
    oracle.render_levelset(torus_small.oracle_handle, cam, sh, ofilm, threads=4)
    for cap, b in ((1, 1), (3, 2), (4096, 4096), (128, 0)):
        ctx.set_tuning(ls_probe_cap=cap, ls_probe_b=b)
        film = refapi.new_film(W, H)
        ctx.render_levelset(g, cam, sh, film, opts=ctx.ls_opts(order=True))
        assert np.array_equal(film, ofilm), (cap, b)
    g.free()


def test_history_ordering_over_a_sequence_of_frames(tuned, oracle, torus_small, sphere100):
    """heavy tiles first from the previous frame's tile costs (k_order_from_history): frame 1 is rendered in tile order and records the
    costs, frames 2.. are ordered -- all are the oracle's frame, for any thresholds, with the long-ray rounds on and off, for several
    samples per pixel; a different partition, film size or grid starts a new history"""
    ctx = tuned
    g = ctx.upload(torus_small.buf)
    g2 = ctx.upload(sphere100.buf)
    W, H = 403, 237
    cam = api.vdb_render_camera(W, H, (0.0, 90.0, 255.0), (0, 0, 0))
    cam2 = api.vdb_render_camera(W, H, (0.0, 0.0, 300.0), (0, 0, 0))
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    want = refapi.new_film(W, H)
    oracle.render_levelset(torus_small.oracle_handle, cam, sh, want, threads=4)
    want2 = refapi.new_film(W, H)
    oracle.render_levelset(sphere100.oracle_handle, cam2, sh, want2, threads=4)
    for a, b in ((250, 105), (100, 50), (1, 1), (100000, 100000)):
        ctx.set_tuning(ls_history=1, ls_hist_a=a, ls_hist_b=b)
        for rounds in (False, True):
            launches = []
            for frame in range(4):
                film = refapi.new_film(W, H)
                ctx.render_levelset(g, cam, sh, film, opts=ctx.ls_opts(rounds=rounds))
                launches.append(ctx.last_kernel_ms()[1])
                assert np.array_equal(film, want), (a, b, rounds, frame)
            assert launches[1] == launches[0] + 1 or launches[0] == launches[1]      # the ordering launch appears from frame 2 on
            # another grid with the same film: new history, then back
            film = refapi.new_film(W, H)
            ctx.render_levelset(g2, cam2, sh, film, opts=ctx.ls_opts(rounds=rounds))
            assert np.array_equal(film, want2)
            # three ranks into one film, twice (each call changes the partition -> each starts a new history)
            film = refapi.new_film(W, H)
            for rep in range(2):
                for r in range(3):
                    ctx.render_levelset(g, cam, sh, film, opts=ctx.ls_opts(rounds=rounds, part=api.partition(r, 3, 32, 16)))
            assert np.array_equal(film, want)
    # several samples per pixel
    want5 = refapi.new_film(W, H)
    oracle.render_levelset(torus_small.oracle_handle, cam, sh, want5, spp=3, jitter=api.jitter_table(1), threads=4)
    ctx.set_tuning(ls_history=1, ls_hist_a=250, ls_hist_b=105)
    for frame in range(3):
        film = refapi.new_film(W, H)
        ctx.render_levelset(g, cam, sh, film, opts=ctx.ls_opts(spp=3, seed=1))
        assert np.array_equal(film, want5)
    g.free(); g2.free()


def test_dense_instantiation_for_frames_with_many_tiles(ctx, oracle, torus_small):
    """from kDenseMinTilesPerSm (800) 8x4 tiles per SM on, one-sample float frames run the 6-CTA instantiation of the render kernel
    (ls_dense, default on): the same film and records as the standard instantiation, and as the oracle on a sample of the pixels"""
    g = ctx.upload(torus_small.buf)
    W, H = 2560, 1600                     # 128 000 tiles: 865 per SM on a B200
    cam = api.vdb_render_camera(W, H, (0.0, 90.0, 255.0), (0, 0, 0))
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    out = {}
    try:
        for dense in (1, 0):
            ctx.set_tuning(ls_dense=800 if dense else 0)
            for want_aux in (True, False):
                film = refapi.new_film(W, H)
                aux = refapi.AuxArrays(W, H)
                ctx.render_levelset(g, cam, sh, film, aux=aux.pod() if want_aux else None)
                out[dense, want_aux] = (film, aux)
    finally:
        ctx.set_tuning(ls_dense=800)
    assert out[1, True][1].hit.sum() > 500000
    assert np.array_equal(out[1, True][0], out[0, True][0]) and np.array_equal(out[1, False][0], out[0, False][0])
    assert np.array_equal(out[1, True][0], out[1, False][0])
    assert_records_equal(out[1, True][1], out[0, True][1])
    # against the oracle: a band of rows through the middle of the frame (a partition of the oracle's render)
    ofilm = refapi.new_film(W, H)
    oaux, _ = oracle.render_levelset(torus_small.oracle_handle, cam, sh, ofilm, aux=True, threads=8, part=api.partition(3, 16, 2560, 100))
    rows = slice(300, 400)
    assert np.array_equal(ofilm[rows], out[1, True][0][rows])
    assert np.array_equal(oaux.hit[rows], out[1, True][1].hit[rows]) and np.array_equal(oaux.ijk[rows], out[1, True][1].ijk[rows])
