"""Halo blocks (DevGrid::halo) are an acceleration structure: the frames must not depend on them.  The same level-set and
fog frames are rendered by a second process with VDBRT_HALO=0 (the stencil / sampler then walks the leaves a cell touches, the
fallback the library also takes when there is no memory for the blocks) and compared bit for bit -- including exp(), which is the
same CUDA routine in both processes.  Both are separately held against the oracle in test_gpu_parity.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

from openvdb_b200 import api, _abi as abi
from tests import refapi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 320, 200

CHILD = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
from tests.test_gpu_halo import render_all
np.savez(sys.argv[2], **render_all())
"""


def render_all():
    ctx = api.Context(0)
    out = {}
    ls = ctx.build_torus(60.0, 25.0)
    fog = ctx.build_fog(ctx.build_sphere(40.0, (3.0, -2.0, 1.0)))
    cam = api.vdb_render_camera(W, H, (30.0, 90.0, 240.0), (0.0, 0.0, 0.0))
    for name, kind in (("diffuse", abi.SHADER_DIFFUSE), ("normal", abi.SHADER_NORMAL)):
        film = refapi.new_film(W, H, (0.1, 0.2, 0.3, 0.4))
        aux = refapi.AuxArrays(W, H)
        pod = aux.pod()
        ctx.render_levelset(ls, cam, api.make_shader(kind), film, aux=pod, bg=(0.1, 0.2, 0.3, 0.4))
        out["ls_" + name] = film
        out["t_" + name] = aux.t_index.copy()
        out["nml_" + name] = aux.nml.copy()
    film = refapi.new_film(W, H)
    ctx.render_levelset(ls, cam, api.make_shader(abi.SHADER_DIFFUSE), film, spp=3, seed=7)
    out["ls_spp3"] = film
    vo = api.vol_opts_default()
    vo.primary_step = 0.5
    film = refapi.new_film(W, H)
    ctx.render_volume(fog, api.vdb_render_camera(W, H, (20.0, 30.0, 130.0), (3.0, -2.0, 1.0)), vo, film)
    out["fog"] = film
    ctx.close()
    return out


def test_frames_do_not_depend_on_the_halo_blocks(tmp_path):
    assert os.environ.get("VDBRT_HALO", "1") != "0"
    mine = render_all()
    path = str(tmp_path / "nohalo.npz")
    env = dict(os.environ, VDBRT_HALO="0")
    r = subprocess.run([sys.executable, "-c", CHILD, ROOT, path], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    other = np.load(path)
    assert (mine["ls_diffuse"][..., :3] != np.float32([0.1, 0.2, 0.3])).any(axis=2).sum() > 5000
    assert (mine["fog"][..., 3] > 0.01).sum() > 5000
    for k, v in mine.items():
        assert np.array_equal(v, other[k]), k
