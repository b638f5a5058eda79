#!/usr/bin/env python3
"""Generates the NanoVDB FILE fixtures under tests/golden/ with the reference's own writer (nanovdb::io::writeGrid(s),
compiled as oracle/_ref/ref_nvdb_io by oracle/Makefile) from grids the reference built (oracle/_ref/libvdbref.so):

  io_files.npz  (numpy-compressed container of four byte strings; the tests unpack them into a temporary directory)
    io_sphere.raw     serialised NanoGrid<float>: createLevelSetSphere(6, (20,20,20), 1, 3), grid name "ls_sphere"
    io_none.nvdb      writeGrid(io_sphere.raw, Codec::NONE)
    io_zip.nvdb       writeGrid(io_sphere.raw, Codec::ZIP)
    io_two_zip.nvdb   writeGrids({sdfToFogVolume of the same sphere named "fog_sphere", io_sphere.raw}, Codec::ZIP)
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import refapi  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
TOOL = os.path.join(ROOT, "oracle", "_ref", "ref_nvdb_io")


def named(buf, name):
    b = np.array(buf, np.uint8, copy=True)
    b[40:296] = 0                                     # GridData::mGridName (NanoVDB.h:1954)
    b[40:40 + len(name)] = np.frombuffer(name, np.uint8)
    return b


def main():
    R = refapi.Ref()
    g = R.sphere(6.0, (20.0, 20.0, 20.0))
    ls = named(R.nanovdb(g), b"ls_sphere")
    fog = named(R.nanovdb(R.fog_from_levelset(g)), b"fog_sphere")
    files = {}
    with tempfile.TemporaryDirectory() as tmp:
        sraw, fraw = os.path.join(tmp, "io_sphere.raw"), os.path.join(tmp, "fog.raw")
        ls.tofile(sraw)
        fog.tofile(fraw)
        subprocess.run([TOOL, "write", os.path.join(tmp, "io_none.nvdb"), "none", sraw], check=True)
        subprocess.run([TOOL, "write", os.path.join(tmp, "io_zip.nvdb"), "zip", sraw], check=True)
        subprocess.run([TOOL, "write", os.path.join(tmp, "io_two_zip.nvdb"), "zip", fraw, sraw], check=True)
        for f in ("io_sphere.raw", "io_none.nvdb", "io_zip.nvdb", "io_two_zip.nvdb"):
            files[f.replace(".", "_")] = np.fromfile(os.path.join(tmp, f), np.uint8)
            print(f, files[f.replace(".", "_")].size)
    np.savez_compressed(os.path.join(OUT, "io_files.npz"), **files)
    print("io_files.npz", os.path.getsize(os.path.join(OUT, "io_files.npz")))


if __name__ == "__main__":
    main()
