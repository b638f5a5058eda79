import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
sys.argv = ['x', 'none']
exec(open('/root/repo/tools/sweep_ls.py').read().split("what = sys.argv")[0])
BASE.update(ls_strip=1, ls_refill=32, ls_order=0, ls_affine=0, ls_tail=48)
which = sys.argv0 if False else None
import os
what = os.environ.get("WHAT", "c2 c4").split()
S = [dict(rounds=False)] + [dict(ls_tail=t) for t in (8, 16, 32, 48, 64, 96, 160)]
if 'c2' in what:
    g = ctx.build_torus(650.0, 325.0)
    W, H = 1920, 1080
    cam = api.vdb_render_camera(W, H, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))
    sweep('c2', g, cam, W, H, S)
    sweep('c2', g, cam, W, H, S, shares=(8,))
    sweep('c2', g, cam, W, H, [S[0], S[3], S[4]], shares=(2, 4))
    g.free()
if 'c4' in what:
    g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0))
    W, H = 3840, 2160
    cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))
    sweep('c4', g, cam, W, H, [S[0], S[2], S[4], S[6]])
    sweep('c4', g, cam, W, H, S, shares=(8,))
    sweep('c4', g, cam, W, H, [S[0], S[3], S[4]], shares=(2, 4))
