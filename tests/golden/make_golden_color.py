#!/usr/bin/env python3
"""Generates tests/golden/color_shaders.npz with the UNMODIFIED reference (oracle/_ref/libvdbref.so):
  ls, color      serialised NanoGrid<float> (createLevelSetSphere(12, (1,2,3), 1, 3)) and NanoGrid<Vec3f> (vdbref_color_grid:
                 voxel size 2, translated, one colour per covered voxel + one tile + background) -- createNanoGrid of the OpenVDB grids
  film_<shader>  tools::rayTrace with MatteShader / NormalShader / PositionShader / DiffuseShader <Vec3SGrid>, 160x120, 1 spp
  scene_nvdb     bytes of a ZIP-compressed .nvdb file holding both grids, named "surface" and "Cd", written by the reference's
                 nanovdb::io::writeGrids (oracle/_ref/ref_nvdb_io) -- the input of vdbrt_render's -color test
"""
import subprocess
import tempfile
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import refapi  # noqa: E402
from tests.test_color_shaders import W, H, KINDS, shader  # noqa: E402


def main():
    R = refapi.Ref()
    ls = R.sphere(12.0, (1.0, 2.0, 3.0))
    col = R.color_grid(ls, 2.0, (0.5, -0.25, 0.0))
    out = {"ls": np.array(R.nanovdb(ls)), "color": np.array(R.color_nanovdb(col))}
    d = refapi.camera_desc(W, H, translation=(18.0, 25.0, 90.0), lookat=(1.0, 2.0, 3.0))
    for name, kind in KINDS:
        film = refapi.new_film(W, H)
        R.render_levelset_color(ls, col, d, shader(kind), film)
        out["film_" + name] = film
        print(name, int((film[..., :3] != 0).any(axis=2).sum()), "coloured pixels")
    def named(buf, name):
        b = np.array(buf, np.uint8, copy=True)
        b[40:296] = 0
        b[40:40 + len(name)] = np.frombuffer(name, np.uint8)
        return b
    with tempfile.TemporaryDirectory() as tmp:
        named(out["ls"], b"surface").tofile(os.path.join(tmp, "ls.raw"))
        named(out["color"], b"Cd").tofile(os.path.join(tmp, "cd.raw"))
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_nvdb_io"), "write", os.path.join(tmp, "scene.nvdb"), "zip",
                        os.path.join(tmp, "ls.raw"), os.path.join(tmp, "cd.raw")], check=True)
        out["scene_nvdb"] = np.fromfile(os.path.join(tmp, "scene.nvdb"), np.uint8)
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "color_shaders.npz")
    np.savez_compressed(p, **out)
    print(p, os.path.getsize(p))


if __name__ == "__main__":
    main()
