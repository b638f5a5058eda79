"""Quantised grids (NanoGrid<Fp4|Fp8|Fp16|FpN>; nanovdb/nanovdb/NanoVDB.h:3752-3980) -- CPU side.

The reference ray tracer works on OpenVDB FloatGrids; a quantised NanoVDB grid reaches it through
nanovdb::tools::nanoToOpenVDB, which reads every voxel with LeafData<FpX>::getValue (float(code) * mQuantum + mMinimum).
Here the oracle port (which dequantises in place, oracle/vdbrt_oracle.cc leafValue) is pinned bit-for-bit against
  * the unmodified reference run in this process: createNanoGrid<FloatGrid, FpX> -> nanoToOpenVDB -> probeValue / rayTrace / VolumeRender
  * the committed fixtures the reference produced (tests/golden/quantized.npz, tests/golden/make_golden_quant.py).
The CUDA path is checked against the same fixtures and the oracle in tests/test_gpu_quantized.py."""
import os

import numpy as np
import pytest

from openvdb_b200 import api, _abi as abi
from tests import refapi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quantized.npz")
W, H = 96, 72
FW, FH = 64, 48
TYPES = [("fp4", 13, -1.0), ("fp8", 14, -1.0), ("fp16", 15, -1.0), ("fpn", 16, -1.0), ("fpn_loose", 16, 0.2)]
CENTER = (1.5, -2.0, 0.5)
EYE = (20.0, 14.0, 60.0)


def golden_sphere(ref):
    return ref.sphere(14.0, CENTER)


def ls_camera():
    return api.vdb_render_camera(W, H, EYE, CENTER), refapi.camera_desc(W, H, translation=EYE, lookat=CENTER)


def fog_camera():
    return api.vdb_render_camera(FW, FH, EYE, CENTER), refapi.camera_desc(FW, FH, translation=EYE, lookat=CENTER)


def fog_opts(defaults):
    o = defaults
    o.primary_step = 0.5
    o.shadow_step = 2.0
    return o


def probe_points():
    return np.random.default_rng(7).integers(-22, 23, size=(30000, 3)).astype(np.int32)


def fpn_widths(buf):
    """histogram of log2(bit width) over the leaves of an FpN grid, and the offset where the last leaf ends"""
    tree = 672
    leaf_off = tree + int(np.frombuffer(buf[tree:tree + 8].tobytes(), np.int64)[0])
    n = int(np.frombuffer(buf[tree + 32:tree + 36].tobytes(), np.uint32)[0])
    hist, off = {}, leaf_off
    for _ in range(n):
        b = int(buf[off + 15]) >> 5
        hist[b] = hist.get(b, 0) + 1
        off += 96 + 64 * (1 << b)
    return hist, off


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("name,gtype,tol", TYPES)
def test_oracle_dequantises_like_the_reference(ref, oracle, name, gtype, tol):
    g = golden_sphere(ref)
    q = ref.nanovdb_quantized(g, gtype, tolerance=tol)
    assert int(np.frombuffer(q[636:640].tobytes(), np.uint32)[0]) == gtype
    og = oracle.open(q)
    assert oracle.info(og).source_type == gtype
    rq = ref.from_nanovdb(q)                      # nanoToOpenVDB: the FloatGrid the reference would ray-trace
    ijk = probe_points()
    rv, ra = ref.probe(rq, ijk)
    ov, oa = oracle.probe(og, ijk)
    assert np.array_equal(rv.view(np.uint32), ov.view(np.uint32)) and np.array_equal(ra, oa)
    if gtype != 15:                               # the quantisation is visible (16 bits are not at float32 print precision)
        fv, _ = ref.probe(g, ijk)
        assert 0 < np.abs(fv - rv).max() < 0.5
    # level-set frame, records and counters
    cam, d = ls_camera()
    assert bytes(cam) == bytes(ref.camera_pod(d))
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    rfilm = refapi.new_film(W, H)
    ref.render_levelset(rq, d, sh, rfilm)
    raux, _, mism = ref.levelset_records(rq, d)
    assert mism == 0
    ofilm = refapi.new_film(W, H)
    oaux, _ = oracle.render_levelset(og, cam, sh, ofilm, aux=True)
    assert raux.hit.sum() > 800
    for k in ("hit", "ijk", "t_index", "t_world", "xyz", "nml"):
        assert np.array_equal(getattr(raux, k), getattr(oaux, k)), k
    assert np.array_equal(rfilm, ofilm)
    ref.free(rq); ref.free(g); oracle.close(og)


@pytest.mark.parametrize("gtype", [13, 14, 16])
def test_dithered_quantisation(ref, oracle, gtype):
    """createNanoGrid(.., ditherOn = true) chooses other codes (DitherLUT); the leaves decode the same way"""
    g = golden_sphere(ref)
    plain = ref.nanovdb_quantized(g, gtype)
    q = ref.nanovdb_quantized(g, gtype, dither=True)
    assert (q.size == plain.size or gtype == 16) and not np.array_equal(q, plain)        # FpN may also pick other bit widths
    og = oracle.open(q)
    rq = ref.from_nanovdb(q)
    ijk = probe_points()
    rv, ra = ref.probe(rq, ijk)
    ov, oa = oracle.probe(og, ijk)
    assert np.array_equal(rv.view(np.uint32), ov.view(np.uint32)) and np.array_equal(ra, oa)
    cam, d = ls_camera()
    sh = api.make_shader(abi.SHADER_NORMAL)
    rfilm = refapi.new_film(W, H)
    ref.render_levelset(rq, d, sh, rfilm)
    ofilm = refapi.new_film(W, H)
    oracle.render_levelset(og, cam, sh, ofilm)
    assert np.array_equal(rfilm, ofilm)
    ref.free(rq); ref.free(g); oracle.close(og)


def test_fpn_fixture_mixes_bit_widths(gold):
    hist, end = fpn_widths(gold["ls_fpn_loose"])
    assert len(hist) >= 3, hist                   # 1-, 2-, 4-bit leaves ...
    hist2, _ = fpn_widths(gold["ls_fpn"])
    assert set(hist) != set(hist2)                # ... and 8-bit ones with the default tolerance
    assert end <= gold["ls_fpn_loose"].size


@pytest.mark.parametrize("name,gtype,tol", [t for t in TYPES if t[0] in ("fp8", "fpn")])
def test_oracle_fog_like_the_reference(ref, oracle, name, gtype, tol):
    ls = golden_sphere(ref)
    fog = ref.fog_from_levelset(ls)
    q = ref.nanovdb_quantized(fog, gtype, tolerance=tol)
    og = oracle.open(q)
    rq = ref.from_nanovdb(q)
    cam, d = fog_camera()
    o = fog_opts(ref.vol_defaults())
    rfilm = refapi.new_film(FW, FH)
    ref.render_volume(rq, d, o, rfilm)
    ofilm = refapi.new_film(FW, FH)
    oracle.render_volume(og, cam, o, ofilm)
    assert (rfilm[..., 3] > 0.01).sum() > 300
    assert np.array_equal(rfilm, ofilm)
    for h in (rq, fog, ls):
        ref.free(h)
    oracle.close(og)


@pytest.mark.parametrize("name", [t[0] for t in TYPES])
def test_oracle_against_golden_fixture(oracle, gold, name):
    """no reference needed: buffers and frames were written by the reference (make_golden_quant.py)"""
    q = refapi.aligned_copy(gold["ls_" + name])
    og = oracle.open(q)
    ov, oa = oracle.probe(og, probe_points())
    assert np.array_equal(ov.view(np.uint32), gold["probe_" + name].view(np.uint32))
    assert np.array_equal(oa, gold["active_" + name])
    cam, _ = ls_camera()
    ofilm = refapi.new_film(W, H)
    oaux, _ = oracle.render_levelset(og, cam, api.make_shader(abi.SHADER_DIFFUSE), ofilm, aux=True)
    assert np.array_equal(ofilm, gold["film_" + name])
    assert np.array_equal(oaux.hit, gold["hit_" + name]) and np.array_equal(oaux.ijk, gold["ijk_" + name])
    oracle.close(og)


@pytest.mark.parametrize("name", ["fp8", "fpn"])
def test_oracle_fog_against_golden_fixture(oracle, gold, name):
    og = oracle.open(refapi.aligned_copy(gold["fog_" + name]))
    cam, _ = fog_camera()
    opts = abi.VolOpts.from_buffer_copy(gold["fog_opts"].tobytes())
    ofilm = refapi.new_film(FW, FH)
    oracle.render_volume(og, cam, opts, ofilm)
    assert np.array_equal(ofilm, gold["fogfilm_" + name])
    oracle.close(og)
