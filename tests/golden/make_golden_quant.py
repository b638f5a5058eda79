#!/usr/bin/env python3
"""Generates tests/golden/quantized.npz with the UNMODIFIED reference (oracle/_ref/libvdbref.so):
  ls_<t>        createNanoGrid<FloatGrid, Fp4|Fp8|Fp16|FpN>(createLevelSetSphere(14, (1.5,-2,0.5), 1, 3)); fpn_loose = FpN with
                AbsDiff(0.2), which mixes 1-, 2- and 4-bit leaves
  probe_<t>, active_<t>   probeValue of nanoToOpenVDB(ls_<t>) at 30 000 seeded coordinates
  film_<t>, hit_<t>, ijk_<t>   tools::rayTrace (DiffuseShader, 96x72) of nanoToOpenVDB(ls_<t>) and the first-hit voxels
  fog_<t>, fogfilm_<t>    sdfToFogVolume of the sphere, quantised (Fp8, FpN), and VolumeRender of nanoToOpenVDB(fog_<t>), 64x48
  fog_opts      the vdbrt_vol_opts bytes used (VolumeRender defaults, step 0.5, shadow step 2)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openvdb_b200 import api, _abi as abi  # noqa: E402
from tests import refapi  # noqa: E402
from tests.test_quantized import W, H, FW, FH, TYPES, golden_sphere, ls_camera, fog_camera, fog_opts, probe_points  # noqa: E402


def main():
    R = refapi.Ref()
    ls = golden_sphere(R)
    fog = R.fog_from_levelset(ls)
    out = {}
    _, d = ls_camera()
    _, fd = fog_camera()
    opts = fog_opts(R.vol_defaults())
    out["fog_opts"] = np.frombuffer(bytes(opts), np.uint8).copy()
    ijk = probe_points()
    for name, gtype, tol in TYPES:
        q = R.nanovdb_quantized(ls, gtype, tolerance=tol)
        out["ls_" + name] = np.array(q)
        rq = R.from_nanovdb(q)
        v, a = R.probe(rq, ijk)
        out["probe_" + name], out["active_" + name] = v, a
        film = refapi.new_film(W, H)
        R.render_levelset(rq, d, api.make_shader(abi.SHADER_DIFFUSE), film)
        aux, _, mism = R.levelset_records(rq, d)
        assert mism == 0
        out["film_" + name], out["hit_" + name], out["ijk_" + name] = film, aux.hit.copy(), aux.ijk.copy()
        print(name, q.size, "bytes,", int(aux.hit.sum()), "hits")
        R.free(rq)
        if name in ("fp8", "fpn"):
            fq = R.nanovdb_quantized(fog, gtype, tolerance=tol)
            out["fog_" + name] = np.array(fq)
            rq = R.from_nanovdb(fq)
            ff = refapi.new_film(FW, FH)
            R.render_volume(rq, fd, opts, ff)
            out["fogfilm_" + name] = ff
            print(" fog", fq.size, "bytes,", int((ff[..., 3] > 0.01).sum()), "covered pixels")
            R.free(rq)
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quantized.npz")
    np.savez_compressed(p, **out)
    print(p, os.path.getsize(p))


if __name__ == "__main__":
    main()
