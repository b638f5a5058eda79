"""Quantised grids on the GPU.  NanoGrid<Fp8|Fp16> are rendered AS THEY ARE by their own kernel instantiations
(codes dequantised where a value is fetched, code halo blocks: csrc/vdbrt_device.cuh "leaf kinds"; knob quant_native, default on);
NanoGrid<Fp4|FpN> -- and Fp8 / Fp16 with quant_native = 0 -- have their leaves expanded to floats on the device at upload
(csrc/vdbrt_quant.cu) and then render like any NanoGrid<float>.  Both ways everything is bit-exact:
against the oracle dequantising in place, against the frames the unmodified reference rendered from nanoToOpenVDB of the same
buffers (tests/golden/quantized.npz) and -- when oracle/_ref travelled -- against the reference run here."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from openvdb_b200 import api, _abi as abi
from tests import refapi
from tests.test_quantized import W, H, FW, FH, TYPES, GOLD, ls_camera, fog_camera, probe_points, fpn_widths
from tests.test_gpu_parity import gpu_levelset, assert_records_equal, RTOL, ATOL

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "openvdb_b200", "vdbrt_render")


def source_leaf_offsets(q, gtype, leaf_off, n):
    if gtype != 16:
        stride = 96 + 64 * {13: 4, 14: 8, 15: 16}[gtype]
        return [leaf_off + stride * i for i in range(n)]
    offs, off = [], leaf_off
    for _ in range(n):
        offs.append(off)
        off += 96 + 64 * (1 << (int(q[off + 15]) >> 5))
    return offs


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(params=[1, 0], ids=["native", "expanded"])
def mode(ctx, request):
    """quant_native = 1: Fp8 / Fp16 rendered from their codes; 0: every quantised grid expanded at upload"""
    ctx.set_tuning(quant_native=request.param)
    yield request.param
    ctx.set_tuning(quant_native=1)


@pytest.mark.parametrize("name,gtype", [(t[0], t[1]) for t in TYPES])
def test_expanded_grid_is_the_dequantised_float_grid(ctx, oracle, gold, name, gtype):
    q = refapi.aligned_copy(gold["ls_" + name])
    og = oracle.open(q)
    ctx.set_tuning(quant_native=0)
    try:
        g = ctx.upload(q)
    finally:
        ctx.set_tuning(quant_native=1)
    assert g.info.leaf_kind == 0
    qi = oracle.info(og)
    assert g.info.source_type == gtype
    assert g.info.leaf_count == qi.leaf_count and g.info.active_voxels == qi.active_voxels
    assert list(g.info.node_bbox) == list(qi.node_bbox) and list(g.info.index_bbox) == list(qi.index_bbox)
    buf = g.download()
    tree = 672
    leaf_off = tree + int(np.frombuffer(q[tree:tree + 8].tobytes(), np.int64)[0])
    assert buf.size == g.info.bytes == leaf_off + 2144 * qi.leaf_count
    assert int(np.frombuffer(buf[636:640].tobytes(), np.uint32)[0]) == 1          # a NanoGrid<float> now
    assert int(np.frombuffer(buf[32:40].tobytes(), np.uint64)[0]) == buf.size
    eg = oracle.open(buf)                                                        # the oracle reads it as a plain float grid
    assert oracle.info(eg).source_type == 1
    ijk = probe_points()
    ev, ea = oracle.probe(eg, ijk)
    assert np.array_equal(ev.view(np.uint32), gold["probe_" + name].view(np.uint32))      # == the reference's nanoToOpenVDB values
    assert np.array_equal(ea, gold["active_" + name])
    # every voxel of every leaf, and the leaf statistics (LeafFnBase::getMin/getMax/getAvg/getDev)
    n = qi.leaf_count
    leaves = buf[leaf_off:leaf_off + 2144 * n].reshape(n, 2144)
    origins = np.ascontiguousarray(leaves[:, :12]).view(np.int32).reshape(n, 3)
    vals = np.ascontiguousarray(leaves[:, 96:]).view(np.float32).reshape(n, 512)
    offs = source_leaf_offsets(q, gtype, leaf_off, n)                             # expanded leaf i = i-th source leaf by address
    x, y, z = np.meshgrid(np.arange(8), np.arange(8), np.arange(8), indexing="ij")
    for li in np.random.default_rng(3).choice(n, size=min(n, 40), replace=False):
        src = q[offs[li]:offs[li] + 96]
        assert np.array_equal(src[:12], leaves[li, :12]) and np.array_equal(src[16:80], leaves[li, 16:80])   # origin, value mask
        o = origins[li] & ~7                                                     # mBBoxMin is the min of the ACTIVE voxels; the origin is its 8-aligned part
        c = np.stack([o[0] + x.ravel(), o[1] + y.ravel(), o[2] + z.ravel()], axis=1).astype(np.int32)
        qv, _ = oracle.probe(og, c)
        assert np.array_equal(qv.view(np.uint32), vals[li].view(np.uint32))
        minimum, quantum = np.ascontiguousarray(src[80:88]).view(np.float32)
        codes = np.ascontiguousarray(src[88:96]).view(np.uint16).astype(np.float32)
        want = np.array([codes[0] * quantum + minimum, codes[1] * quantum + minimum, codes[2] * quantum + minimum, codes[3] * quantum], np.float32)
        assert np.array_equal(np.ascontiguousarray(leaves[li, 80:96]).view(np.float32).view(np.uint32), want.view(np.uint32))
    assert (leaves[:, 15] >> 5 == 0).all()                                       # no FpN bit width left in mFlags
    g.free(); oracle.close(og); oracle.close(eg)


@pytest.mark.parametrize("name,gtype", [(t[0], t[1]) for t in TYPES])
def test_levelset_render_of_quantised_grid(ctx, oracle, gold, mode, name, gtype):
    q = refapi.aligned_copy(gold["ls_" + name])
    og = oracle.open(q)
    g = ctx.upload(q)
    assert g.info.leaf_kind == ({14: 1, 15: 2}.get(gtype, 0) if mode else 0)
    cam, _ = ls_camera()
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    film, aux = gpu_levelset(ctx, g, cam, sh, W, H)
    ofilm = refapi.new_film(W, H)
    oaux, _ = oracle.render_levelset(og, cam, sh, ofilm, aux=True)
    assert_records_equal(aux, oaux)
    assert np.array_equal(film, ofilm)
    assert np.array_equal(film, gold["film_" + name])                            # the reference's frame
    assert np.array_equal(aux.hit, gold["hit_" + name]) and np.array_equal(aux.ijk, gold["ijk_" + name])
    # supersampled + normal shader against the oracle
    sh2 = api.make_shader(abi.SHADER_NORMAL)
    f2 = refapi.new_film(W, H)
    ctx.render_levelset(g, cam, sh2, f2, spp=4, seed=0)
    o2 = refapi.new_film(W, H)
    oracle.render_levelset(og, cam, sh2, o2, spp=4, jitter=api.jitter_table(0))
    assert np.array_equal(f2, o2)
    g.free(); oracle.close(og)


@pytest.mark.parametrize("name", ["fp8", "fpn"])
def test_fog_render_of_quantised_grid(ctx, oracle, gold, mode, name):
    q = refapi.aligned_copy(gold["fog_" + name])
    g = ctx.upload(q)
    # both fog paths: the wavefront (default) and the one-loop kernel
    cam0, _ = fog_camera()
    opts0 = abi.VolOpts.from_buffer_copy(gold["fog_opts"].tobytes())
    ctx.set_tuning(fog_wave=0)
    one = refapi.new_film(FW, FH)
    ctx.render_volume(g, cam0, opts0, one)
    ctx.set_tuning(fog_wave=1)
    two = refapi.new_film(FW, FH)
    ctx.render_volume(g, cam0, opts0, two)
    assert np.array_equal(one, two)
    assert g.info.grid_class == abi.GRID_CLASS_FOG_VOLUME and g.info.source_type in (14, 16)
    cam, _ = fog_camera()
    opts = abi.VolOpts.from_buffer_copy(gold["fog_opts"].tobytes())
    film = refapi.new_film(FW, FH)
    ctx.render_volume(g, cam, opts, film)
    want = gold["fogfilm_" + name]                                               # the reference's VolumeRender of nanoToOpenVDB(q)
    assert (want[..., 3] > 0.01).sum() > 300
    assert np.allclose(film, want, rtol=RTOL, atol=ATOL)
    assert np.abs(film - want).max() < 2e-6                                      # exp() is CUDA's, everything else is bit-exact
    g.free()


def test_quantised_upload_from_device_memory(ctx, oracle, gold):
    q = refapi.aligned_copy(gold["ls_fpn_loose"])
    p = C.c_void_p()
    api._check(ctx.L.vdbrt_device_alloc(ctx.handle, q.size, C.byref(p)))
    api.memcpy(ctx, p, q.ctypes.data, q.size, 0)
    ctx.synchronize()
    g = ctx.upload_device(p, q.size)
    h = ctx.upload(q)
    assert g.info.source_type == 16 and np.array_equal(g.download(), h.download())
    api._check(ctx.L.vdbrt_device_free(ctx.handle, p))
    g.free(); h.free()


def test_corrupt_quantised_grids_are_rejected(ctx, gold, mode):
    q = refapi.aligned_copy(gold["ls_fp8"])
    tree = 672
    lower_off = tree + int(np.frombuffer(q[tree + 8:tree + 16].tobytes(), np.int64)[0])
    cmask = np.ascontiguousarray(q[lower_off + 32 + 512:lower_off + 32 + 1024]).view(np.uint64)
    word = int(np.flatnonzero(cmask)[0])
    bit = int(cmask[word]).bit_length() - 1
    slot = 64 * word + bit
    bad = q.copy()
    bad[lower_off + 1088 + 8 * slot:lower_off + 1088 + 8 * slot + 8] = np.frombuffer(np.int64(1 << 40).tobytes(), np.uint8)
    with pytest.raises(api.VdbrtError) as e:
        ctx.upload(bad)
    assert e.value.code == abi.ERR_BAD_GRID and "child offset" in str(e.value)
    bad = q.copy()                                                               # one child bit too many
    free = int(np.flatnonzero(cmask != np.uint64(0xFFFFFFFFFFFFFFFF))[0])
    fb = [b for b in range(64) if not (int(cmask[free]) >> b) & 1][0]
    w = np.uint64(int(cmask[free]) | (1 << fb))
    bad[lower_off + 32 + 512 + 8 * free:lower_off + 32 + 512 + 8 * free + 8] = np.frombuffer(w.tobytes(), np.uint8)
    with pytest.raises(api.VdbrtError) as e:
        ctx.upload(bad)
    assert e.value.code == abi.ERR_BAD_GRID
    fpn = refapi.aligned_copy(gold["ls_fpn"])                                    # an FpN leaf that claims 32-bit codes
    leaf_off = tree + int(np.frombuffer(fpn[tree:tree + 8].tobytes(), np.int64)[0])
    fpn[leaf_off + 15] = (fpn[leaf_off + 15] & 0x1F) | (5 << 5)
    with pytest.raises(api.VdbrtError) as e:
        ctx.upload(fpn)
    assert e.value.code in (abi.ERR_UNSUPPORTED, abi.ERR_BAD_GRID)
    ok = ctx.upload(q)                                                           # the context is still usable
    assert ok.info.source_type == 14
    ok.free()


def test_quantised_grid_against_the_reference_directly(ctx, ref, mode):
    """a larger sphere than the fixture, all four types, reference run here: createNanoGrid<.., FpX> -> upload / nanoToOpenVDB -> rayTrace"""
    ls = ref.sphere(60.0, (3.0, -1.0, 2.0))
    Wd, Hd = 200, 160
    d = refapi.camera_desc(Wd, Hd, translation=(40.0, 50.0, 190.0), lookat=(3.0, -1.0, 2.0))
    cam = api.vdb_render_camera(Wd, Hd, (40.0, 50.0, 190.0), (3.0, -1.0, 2.0))
    sh = api.make_shader(abi.SHADER_DIFFUSE)
    for gtype, tol in ((13, -1.0), (14, -1.0), (15, -1.0), (16, -1.0), (16, 0.1)):
        q = ref.nanovdb_quantized(ls, gtype, tolerance=tol)
        rq = ref.from_nanovdb(q)
        g = ctx.upload(q)
        film, aux = gpu_levelset(ctx, g, cam, sh, Wd, Hd)
        rfilm = refapi.new_film(Wd, Hd)
        ref.render_levelset(rq, d, sh, rfilm, threaded=True)
        raux, _, mism = ref.levelset_records(rq, d)
        assert mism == 0 and raux.hit.sum() > 5000
        assert_records_equal(aux, raux)
        assert np.array_equal(film, rfilm), (gtype, tol)
        g.free(); ref.free(rq)
    ref.free(ls)


def test_command_line_takes_a_quantised_file(ctx, oracle, gold, tmp_path):
    """vdb_render's "first floating-point volume" rule includes the quantised float types"""
    q = refapi.aligned_copy(gold["ls_fp8"])
    path = tmp_path / "fp8.nvdb"
    api.nvdb_write(str(path), q, abi.CODEC_ZIP)
    meta = api.nvdb_list(str(path))
    assert len(meta) == 1 and meta[0].grid_type == 14
    back = api.nvdb_read(str(path))
    assert np.array_equal(back, q)
    out = tmp_path / "fp8.ppm"
    r = subprocess.run([CLI, str(path), str(out), "-res", "%dx%d" % (W, H), "-translate", "20,14,60", "-lookat", "1.5,-2,0.5"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    from tests.test_gpu_cli import read_ppm, to_bits
    assert np.array_equal(read_ppm(str(out)), to_bits(gold["film_fp8"]))


def test_native_quantised_grid_memory_and_limits(ctx, ref, oracle):
    """what rendering from the codes buys: the resident grid (buffer + halo blocks + masks) of an Fp8 / Fp16 grid against the float
    one; arbitrary rays, spans and the things a native grid does not do (work counters, search iterations, sdfToFogVolume)"""
    ls = ref.sphere(60.0, (3.0, -1.0, 2.0))
    fbuf = ref.nanovdb(ls)
    gf = ctx.upload(fbuf)
    sizes = {"float": gf.info.resident_bytes}
    rng = np.random.default_rng(9)
    n = 4000
    eyes = np.column_stack([rng.uniform(-70, 70, n), rng.uniform(-70, 70, n), np.full(n, 200.0)])
    dirs = np.column_stack([rng.uniform(-0.1, 0.1, n), rng.uniform(-0.1, 0.1, n), np.full(n, -1.0)])
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    rays = refapi.make_rays(eyes, dirs)
    for gtype, kind in ((14, 1), (15, 2)):
        q = ref.nanovdb_quantized(ls, gtype)
        g = ctx.upload(q)
        assert g.info.leaf_kind == kind and g.info.source_type == gtype and g.info.bytes == q.size
        assert np.array_equal(g.download(), q)                                   # the buffer is resident as it was uploaded
        sizes[gtype] = g.info.resident_bytes
        og = oracle.open(q)
        got = refapi.hits_to_dict(ctx.intersect(g, rays), n)
        want = oracle.intersect(og, rays)
        assert want["hit"].sum() > 1000 and got.tobytes() == want.tobytes()
        cam = api.vdb_render_camera(64, 48, (40.0, 50.0, 190.0), (3.0, -1.0, 2.0))
        with pytest.raises(api.VdbrtError) as e:
            ctx.count_levelset(g, cam)
        assert e.value.code == abi.ERR_UNSUPPORTED
        with pytest.raises(api.VdbrtError) as e:
            ctx.render_levelset(g, cam, api.make_shader(), refapi.new_film(64, 48), opts=ctx.ls_opts(iterations=2))
        assert e.value.code == abi.ERR_UNSUPPORTED
        with pytest.raises(api.VdbrtError) as e:
            ctx.build_fog(g)
        assert e.value.code == abi.ERR_UNSUPPORTED
        g.free(); oracle.close(og)
    print("resident bytes: float %d, Fp8 %d (%.2fx), Fp16 %d (%.2fx)" % (sizes["float"], sizes[14], sizes["float"] / sizes[14], sizes[15], sizes["float"] / sizes[15]))
    assert sizes[14] < 0.5 * sizes["float"] and sizes[15] < 0.7 * sizes["float"]      # (a small grid: its internal nodes weigh as much as its leaves)
    gf.free(); ref.free(ls)
