import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
sys.argv = ['x', 'none']
exec(open('/root/repo/tools/sweep_ls.py').read().split("what = sys.argv")[0])
BASE.clear(); BASE.update(ls_history=1)
L = [dict(rounds=False), dict(rounds=True)]
g = ctx.build_torus(650.0, 325.0)
W, H = 1920, 1080
cam = api.vdb_render_camera(W, H, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))
sweep('c2', g, cam, W, H, L, shares=(2, 4))
sweep('c2', g, cam, W, H, L)
g.free()
g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0))
W, H = 3840, 2160
cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))
sweep('c4', g, cam, W, H, L, shares=(8,))
sweep('c4', g, cam, W, H, L)
