#!/usr/bin/env python3
"""Generates the committed golden fixtures under tests/golden/ by running the UNMODIFIED reference
(oracle/_ref/libvdbref.so, built from /root/reference by oracle/Makefile) in this container.

  jitter_seed0_seed7.npy   math::Rand01<double>(seed) tables (tools/RayTracer.h:883-885)
  ls_sphere40.npz          tools::rayTrace + LevelSetRayIntersector records on createLevelSetSphere(40,(3,-2,1),1,3), 96x72
  fog_sphere40.npz         VolumeRender on sdfToFogVolume of the same sphere, 64x48, step 0.5
  spans_kat.npz            VolumeRayIntersector::hits spans for the TestVolumeRayIntersector grids

The GPU box has no /root/reference: the -m gpu tests compare against these files (and against the oracle port).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import refapi  # noqa: E402
from openvdb_b200 import _abi as abi  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    R = refapi.Ref()
    np.save(os.path.join(OUT, "jitter_seed0_seed7.npy"), np.stack([R.jitter_table(0), R.jitter_table(7)]))

    g = R.sphere(40.0, (3.0, -2.0, 1.0))
    W, H = 96, 72
    d = refapi.camera_desc(W, H, translation=(30.0, 20.0, 140.0), lookat=(0, 0, 0))
    out = {}
    for name, kind in (("diffuse", abi.SHADER_DIFFUSE), ("normal", abi.SHADER_NORMAL), ("matte", abi.SHADER_MATTE)):
        film = refapi.new_film(W, H, (0.1, 0.2, 0.3, 0.5))
        R.render_levelset(g, d, refapi.shader(kind, (0.9, 0.8, 0.7, 1.0)), film)
        out["film_" + name] = film
    film = refapi.new_film(W, H)
    R.render_levelset(g, d, refapi.shader(abi.SHADER_DIFFUSE), film, spp=4, seed=0, threaded=False)
    out["film_spp4"] = film
    aux, ctr, mism = R.levelset_records(g, d)
    assert mism == 0
    for k in ("hit", "ijk", "t_index", "t_world", "xyz", "nml"):
        out[k] = getattr(aux, k)
    np.savez_compressed(os.path.join(OUT, "ls_sphere40.npz"), **out)

    fg = R.fog_from_levelset(g)
    W, H = 64, 48
    d = refapi.camera_desc(W, H, translation=(30.0, 20.0, 140.0), lookat=(0, 0, 0))
    vo = R.vol_defaults()
    vo.primary_step = 0.5
    film = refapi.new_film(W, H)
    R.render_volume(fg, d, vo, film)
    rays = R.camera_rays(d, [(i, j) for j in range(0, H, 4) for i in range(0, W, 4)])
    spans, counts = R.volume_spans(fg, rays)
    np.savez_compressed(os.path.join(OUT, "fog_sphere40.npz"), film=film, spans=spans, counts=counts)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
