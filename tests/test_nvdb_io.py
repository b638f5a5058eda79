"""NanoVDB file ingestion (openvdb_b200/csrc/vdbrt_io.cc, host-only) against the reference's own reader / writer
(nanovdb/io/IO.h compiled as oracle/_ref/ref_nvdb_io) and against committed files the reference wrote (tests/golden/)."""
import os
import subprocess

import numpy as np
import pytest

from openvdb_b200 import api, _abi as abi

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
TOOL = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "ref_nvdb_io")


_unpacked = {}


def golden(name):
    """the reference-written files live in one compressed container (tests/golden/io_files.npz, make_golden_io.py)"""
    if not _unpacked:
        import tempfile
        d = tempfile.mkdtemp(prefix="vdbrt_io_")
        z = np.load(os.path.join(GOLD, "io_files.npz"))
        for k in z.files:
            base, ext = k.rsplit("_", 1)
            z[k].tofile(os.path.join(d, base + "." + ext))
            _unpacked[base + "." + ext] = os.path.join(d, base + "." + ext)
    return _unpacked[name]


def test_reads_reference_written_files():
    """tests/golden/io_*.nvdb were written by nanovdb::io::writeGrid(s) (make_golden_io.py); io_sphere.raw is the grid buffer"""
    want = np.fromfile(golden("io_sphere.raw"), np.uint8)
    for f in ("io_none.nvdb", "io_zip.nvdb"):
        got = api.nvdb_read(golden(f))
        assert got.ctypes.data % 32 == 0
        assert np.array_equal(got, want), f
        meta = api.nvdb_list(golden(f))
        assert len(meta) == 1 and meta[0].grid_type == 1 and meta[0].grid_class == abi.GRID_CLASS_LEVEL_SET
        assert meta[0].grid_bytes == want.size and meta[0].name == b"ls_sphere"
        assert meta[0].codec == (abi.CODEC_ZIP if "zip" in f else abi.CODEC_NONE)
    # a raw grid buffer is accepted as a file too
    assert np.array_equal(api.nvdb_read(golden("io_sphere.raw")), want)
    assert api.nvdb_list(golden("io_sphere.raw"))[0].grid_bytes == want.size


def test_multi_grid_file_and_lookup_by_name():
    """two grids in one ZIP segment: the fog comes first, the level set is found by name"""
    meta = api.nvdb_list(golden("io_two_zip.nvdb"))
    assert [m.name for m in meta] == [b"fog_sphere", b"ls_sphere"]
    assert [m.grid_class for m in meta] == [abi.GRID_CLASS_FOG_VOLUME, abi.GRID_CLASS_LEVEL_SET]
    ls = api.nvdb_read(golden("io_two_zip.nvdb"), "ls_sphere")
    assert np.array_equal(ls, np.fromfile(golden("io_sphere.raw"), np.uint8))
    first = api.nvdb_read(golden("io_two_zip.nvdb"))          # vdb_render: the first float grid
    assert bytes(first[40:50]) == b"fog_sphere"
    with pytest.raises(api.VdbrtError) as e:
        api.nvdb_read(golden("io_two_zip.nvdb"), "nope")
    assert e.value.code == abi.ERR_IO


def test_write_then_read_round_trip(tmp_path):
    want = np.fromfile(golden("io_sphere.raw"), np.uint8)
    for codec in (abi.CODEC_NONE, abi.CODEC_ZIP):
        p = str(tmp_path / ("rt%d.nvdb" % codec))
        api.nvdb_write(p, want, codec)
        assert np.array_equal(api.nvdb_read(p), want)
    # uncompressed output is byte-identical to what the reference wrote
    api.nvdb_write(str(tmp_path / "same.nvdb"), want, abi.CODEC_NONE)
    assert open(str(tmp_path / "same.nvdb"), "rb").read() == open(golden("io_none.nvdb"), "rb").read()


def test_errors(tmp_path):
    with pytest.raises(api.VdbrtError) as e:
        api.nvdb_read(str(tmp_path / "missing.nvdb"))
    assert e.value.code == abi.ERR_IO
    bad = tmp_path / "bad.nvdb"
    bad.write_bytes(b" BDV" + b"\0" * 60)               # an OpenVDB file starts with 0x56444220
    with pytest.raises(api.VdbrtError) as e:
        api.nvdb_read(str(bad))
    assert e.value.code == abi.ERR_BAD_GRID and "OpenVDB file" in str(e.value)
    trunc = tmp_path / "trunc.nvdb"
    trunc.write_bytes(open(golden("io_none.nvdb"), "rb").read()[:5000])
    with pytest.raises(api.VdbrtError):
        api.nvdb_read(str(trunc))


@pytest.mark.skipif(not os.path.exists(TOOL), reason="oracle/_ref/ref_nvdb_io not built (needs /root/reference)")
def test_against_the_reference_tool(tmp_path):
    """both directions, live: the reference reads what we write, we read what the reference writes"""
    want = np.fromfile(golden("io_sphere.raw"), np.uint8)
    for codec, cname in ((abi.CODEC_NONE, "none"), (abi.CODEC_ZIP, "zip")):
        ours = str(tmp_path / ("ours_%s.nvdb" % cname))
        api.nvdb_write(ours, want, codec)
        back = str(tmp_path / ("back_%s.raw" % cname))
        subprocess.run([TOOL, "read", ours, "ls_sphere", back], check=True)
        assert np.array_equal(np.fromfile(back, np.uint8), want)
        theirs = str(tmp_path / ("theirs_%s.nvdb" % cname))
        subprocess.run([TOOL, "write", theirs, cname, golden("io_sphere.raw")], check=True)
        assert np.array_equal(api.nvdb_read(theirs), want)


def test_save_ppm(tmp_path):
    """Film::savePPM: P6, channel = (unsigned char)(255.0f * v), alpha dropped, '.ppm' appended to a bare name"""
    rng = np.random.default_rng(3)
    film = rng.random((5, 7, 4)).astype(np.float32)
    film[0, 0] = (1.0, 0.0, 0.999999, 0.5)
    api.film_save_ppm(str(tmp_path / "img"), film)
    raw = open(str(tmp_path / "img.ppm"), "rb").read()
    head = b"P6\n7 5\n255\n"
    assert raw.startswith(head)
    px = np.frombuffer(raw[len(head):], np.uint8).reshape(5, 7, 3)
    assert np.array_equal(px, (np.float32(255.0) * film[..., :3]).astype(np.uint8))
    assert tuple(px[0, 0]) == (255, 0, 254)


def _patched(src, tmp_path, name, edits):
    """a copy of a reference-written file with some header bytes overwritten: edits = [(offset, bytes)]"""
    raw = bytearray(open(src, "rb").read())
    for off, b in edits:
        raw[off:off + len(b)] = b
    p = str(tmp_path / name)
    open(p, "wb").write(bytes(raw))
    return p


def test_malformed_files_are_rejected_not_trusted(tmp_path):
    """every size a file states is checked against the length of the file before anything is allocated or read:
    FileHeader 16 B, then FileMetaData 176 B with gridSize @0, fileSize @8, nameSize @136, codec @168"""
    import struct
    none, zipf = golden("io_none.nvdb"), golden("io_zip.nvdb")
    meta = 16
    cases = [
        ("name_wraps.nvdb", none, [(meta + 136, struct.pack("<I", 0xFFFFFFFF))]),          # nameSize + 1 wrapped to 0 in round 1
        ("name_huge.nvdb", none, [(meta + 136, struct.pack("<I", 1 << 20))]),
        ("payload_beyond_eof.nvdb", none, [(meta + 8, struct.pack("<Q", 1 << 40))]),
        ("grid_size_mismatch.nvdb", none, [(meta + 0, struct.pack("<Q", 1 << 50))]),
        ("zip_grid_size_huge.nvdb", zipf, [(meta + 0, struct.pack("<Q", 1 << 62))]),
        ("zip_payload_tiny.nvdb", zipf, [(meta + 8, struct.pack("<Q", 4))]),
    ]
    for name, src, edits in cases:
        p = _patched(src, tmp_path, name, edits)
        for call in (lambda: api.nvdb_read(p), lambda: api.nvdb_list(p)):
            with pytest.raises(api.VdbrtError) as e:
                call()
            assert e.value.code in (abi.ERR_BAD_GRID, abi.ERR_IO), name
    # a ZIP stream that claims to be longer than its payload (the uint64 in front of the stream)
    raw = open(zipf, "rb").read()
    name_size = struct.unpack_from("<I", raw, meta + 136)[0]
    p = _patched(zipf, tmp_path, "zip_stream_long.nvdb", [(16 + 176 + name_size, struct.pack("<Q", 1 << 40))])
    with pytest.raises(api.VdbrtError) as e:
        api.nvdb_read(p)
    assert e.value.code == abi.ERR_BAD_GRID
    # truncated file
    p = str(tmp_path / "cut.nvdb")
    open(p, "wb").write(raw[:len(raw) // 2])
    with pytest.raises(api.VdbrtError):
        api.nvdb_read(p)


def test_raw_grid_buffer_honours_name_and_type(tmp_path):
    raw = golden("io_sphere.raw")
    got = api.nvdb_read(raw)
    assert np.array_equal(got, np.fromfile(raw, np.uint8))
    name = bytes(got[40:40 + 256]).split(b"\0")[0].decode()
    assert np.array_equal(api.nvdb_read(raw, name=name), got)
    with pytest.raises(api.VdbrtError) as e:
        api.nvdb_read(raw, name="no_such_grid")
    assert e.value.code == abi.ERR_IO
    with pytest.raises(api.VdbrtError) as e:
        api.nvdb_read(raw, grid_type=6)                      # a float grid is not a Vec3f colour grid
    assert e.value.code == abi.ERR_NOT_FLOAT
    cut = str(tmp_path / "cut.raw")
    open(cut, "wb").write(open(raw, "rb").read()[:5000])
    with pytest.raises(api.VdbrtError) as e:
        api.nvdb_read(cut)
    assert e.value.code == abi.ERR_BAD_GRID
