"""vdbrt_render (openvdb_b200/csrc/vdbrt_render.cc): the option set of OpenVDB's vdb_render (openvdb_cmd/vdb_render/main.cc)
over the GPU path.  The image it writes must be the PPM of the film the oracle renders with vdb_render's camera rules."""
import os
import subprocess

import numpy as np
import pytest

from openvdb_b200 import api, _abi as abi
from tests import refapi
from tests.test_nvdb_io import golden

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "openvdb_b200", "vdbrt_render")


def read_ppm(path):
    raw = open(path, "rb").read()
    assert raw[:3] == b"P6\n"
    w, h = [int(x) for x in raw.split(b"\n")[1].split()]
    head = len(b"P6\n%d %d\n255\n" % (w, h))
    return np.frombuffer(raw[head:], np.uint8).reshape(h, w, 3)


def read_png(path):
    """8-bit RGB, non-interlaced PNG with filter type 0 rows, decoded with zlib alone (what vdbrt_render writes)"""
    import struct, zlib
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(raw):
        n, kind = struct.unpack(">I4s", raw[pos:pos + 8])
        data = raw[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(kind + data)
        if kind == b"IHDR":
            w, h, depth, colour, comp, filt, lace = struct.unpack(">IIBBBBB", data)
            assert (depth, colour, comp, filt, lace) == (8, 2, 0, 0, 0)
        elif kind == b"IDAT":
            idat += data
        pos += 12 + n
    assert kind == b"IEND"
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 3 * w + 1)
    assert not rows[:, 0].any()
    return rows[:, 1:].reshape(h, w, 3)


def to_bits(film):
    return (np.float32(255.0) * film[..., :3]).astype(np.uint8)          # Film::convertToBitBuffer (tools/RayTracer.h:300-317)


def run(*args):
    return subprocess.run([CLI] + [str(a) for a in args], capture_output=True, text=True, timeout=300)


def test_levelset_from_a_zip_file_with_default_lookat(ctx, oracle, tmp_path):
    """no -rotate / -lookat: the camera points at the centre of the active voxel bbox (main.cc:807-813)"""
    buf = api.nvdb_read(golden("io_zip.nvdb"))
    og = oracle.open(buf)
    info = ctx.upload(buf).info
    centre = [0.5 * (info.index_bbox[a] + info.index_bbox[3 + a]) * info.voxel_size[a] + info.translation[a] for a in range(3)]
    W, H = 200, 120
    out = tmp_path / "ls.ppm"
    r = run(golden("io_zip.nvdb"), out, "-res", "%dx%d" % (W, H), "-translate", "22,25,60", "-v")
    assert r.returncode == 0, r.stderr
    assert "ray-tracing..." in r.stdout and "...completed in" in r.stdout and " -lookat %g,%g,%g" % tuple(centre) in r.stdout
    cam = api.vdb_render_camera(W, H, (22.0, 25.0, 60.0), tuple(centre))
    film = refapi.new_film(W, H)
    oracle.render_levelset(og, cam, api.make_shader(abi.SHADER_DIFFUSE), film, threads=4)
    assert (film[..., 0] > 0).sum() > 500
    assert np.array_equal(read_ppm(str(out)), to_bits(film))


def test_named_grid_normal_shader_samples_and_fov(ctx, oracle, tmp_path):
    buf = api.nvdb_read(golden("io_two_zip.nvdb"), "ls_sphere")
    og = oracle.open(buf)
    W, H = 160, 100
    out = tmp_path / "n.ppm"
    r = run(golden("io_two_zip.nvdb"), out, "-name", "ls_sphere", "-res", "%dx%d" % (W, H), "-t", "20,20,50", "-lookat", "20,20,20",
            "-shader", "normal", "-samples", "4", "-fov", "40")
    assert r.returncode == 0, r.stderr
    # fieldOfViewToFocalLength in double on the float options, stored back as float (main.cc:728-731, RayTracer.h:472-475)
    focal = float(np.float32(float(np.float32(41.2136)) / (2.0 * np.tan(float(np.float32(40.0)) * np.pi / 360.0))))
    cam = api.vdb_render_camera(W, H, (20.0, 20.0, 50.0), (20.0, 20.0, 20.0), focal=focal)
    film = refapi.new_film(W, H)
    oracle.render_levelset(og, cam, api.make_shader(abi.SHADER_NORMAL), film, spp=4, jitter=api.jitter_table(0), threads=1)
    assert np.array_equal(read_ppm(str(out)), to_bits(film))


def test_fog_volume_from_the_first_float_grid(ctx, oracle, tmp_path):
    """io_two_zip.nvdb: the first float grid is the fog volume -> VolumeRender with vdb_render's defaults, -step 0.5"""
    buf = api.nvdb_read(golden("io_two_zip.nvdb"))
    og = oracle.open(buf)
    W, H = 120, 80
    out = tmp_path / "fog.ppm"
    r = run(golden("io_two_zip.nvdb"), out, "-res", "%dx%d" % (W, H), "-t", "20,20,50", "-lookat", "20,20,20", "-step", "0.5",
            "-absorb", "0.4,0.2,0.1", "-light", "0.2,0.5,0.1")
    assert r.returncode == 0, r.stderr
    cam = api.vdb_render_camera(W, H, (20.0, 20.0, 50.0), (20.0, 20.0, 20.0))
    vo = api.vol_opts_default()
    vo.primary_step = 0.5
    n = np.sqrt(0.2 ** 2 + 0.5 ** 2 + 0.1 ** 2)
    for a, v in enumerate((0.2, 0.5, 0.1)):
        vo.light_dir[a] = v / n
    for a, v in enumerate((0.4, 0.2, 0.1)):
        vo.absorption[a] = v
    film = refapi.new_film(W, H)
    oracle.render_volume(og, cam, vo, film, threads=4)
    assert (film[..., 3] > 0).sum() > 300
    got, want = read_ppm(str(out)).astype(int), to_bits(film).astype(int)
    assert np.abs(got - want).max() <= 1           # fog colours are double exp() results: 1e-4 relative -> at most one 8-bit step


def test_color_option(oracle, tmp_path):
    """-color Cd: the Vec3f grid named Cd of the same file colours the surface (main.cc:788-795), here through the position shader
    whose bbox is the world bbox of the active voxels (main.cc:452-458)"""
    z = np.load(os.path.join(ROOT, "tests", "golden", "color_shaders.npz"))
    scene = tmp_path / "scene.nvdb"
    z["scene_nvdb"].tofile(str(scene))
    assert [m.name for m in api.nvdb_list(str(scene))] == [b"surface", b"Cd"]
    ls, col = api.nvdb_read(str(scene), "surface"), api.nvdb_read(str(scene), "Cd", grid_type=6)
    assert np.array_equal(col[296:], refapi.aligned_copy(z["color"])[296:])
    og, oc = oracle.open(ls), oracle.open_color(col)
    W, H = 160, 120
    for name, kind in (("diffuse", abi.SHADER_DIFFUSE), ("position", abi.SHADER_POSITION)):
        out = tmp_path / (name + ".ppm")
        r = run(scene, out, "-res", "%dx%d" % (W, H), "-t", "18,25,90", "-lookat", "1,2,3", "-color", "Cd", "-shader", name)
        assert r.returncode == 0, r.stderr
        info = oracle.info(og)
        lo = [float(info.index_bbox[a]) for a in range(3)]
        hi = [float(info.index_bbox[3 + a]) for a in range(3)]
        sh = api.make_shader(kind, bbox_min=lo, inv_dim=[1.0 / (h - l) for l, h in zip(lo, hi)])
        film = refapi.new_film(W, H)
        oracle.render_levelset(og, api.vdb_render_camera(W, H, (18.0, 25.0, 90.0), (1.0, 2.0, 3.0)), sh, film, color=oc)
        assert np.array_equal(read_ppm(str(out)), to_bits(film)), name
    assert run(scene, tmp_path / "x.ppm", "-color", "surface").returncode != 0          # "surface is not a vec3s color volume"
    assert run(scene, tmp_path / "x.ppm", "-color", "nope").returncode != 0


def test_generators_and_errors(tmp_path):
    r = run("sphere:30", tmp_path / "s.ppm", "-res", "96x64", "-t", "0,0,100")
    assert r.returncode == 0, r.stderr
    img = read_ppm(str(tmp_path / "s.ppm"))
    assert img.shape == (64, 96, 3) and (img[..., 0] > 0).sum() > 500
    assert run("sphere:30", tmp_path / "s.exr").returncode != 0
    assert run("sphere:30", tmp_path / "s.ppm", "-color", "Cd").returncode != 0          # a generator has no file to read Cd from
    assert run("sphere:30", tmp_path / "s.ppm", "-isovalue", "5").returncode != 0          # outside the narrow band -> ValueError
    assert run("sphere:30", tmp_path / "s.ppm", "-samples", "0").returncode != 0
    assert run(tmp_path / "missing.nvdb", tmp_path / "s.ppm").returncode != 0
    assert run("sphere:30", tmp_path / "s.ppm", "-bogus").returncode != 0


def test_png_output_has_the_pixels_of_the_ppm(ctx, oracle, tmp_path):
    """-o x.png: PngWriter (main.cc:335-390) writes Film::convertToBitBuffer<uint8_t>(alpha=false) as 8-bit RGB; .exr is refused the way a
    vdb_render built without OpenEXR refuses it (main.cc:248-253), other extensions the way isExtensionSupported does"""
    W, H = 150, 90
    png, ppm = tmp_path / "s.png", tmp_path / "s.ppm"
    for out in (png, ppm):
        r = run("sphere:40", out, "-res", "%dx%d" % (W, H), "-translate", "30,20,140", "-lookat", "0,0,0", "-shader", "normal")
        assert r.returncode == 0, r.stderr
    a, b = read_png(str(png)), read_ppm(str(ppm))
    assert a.shape == (H, W, 3) and (b.sum(axis=2) > 0).sum() > 1000
    assert np.array_equal(a, b)
    r = run("sphere:40", tmp_path / "s.exr", "-res", "32x32")
    assert r.returncode != 0 and "has not been compiled with .exr support" in r.stderr
    r = run("sphere:40", tmp_path / "s.jpg", "-res", "32x32")
    assert r.returncode != 0 and "unsupported image file format" in r.stderr
