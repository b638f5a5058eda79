import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
sys.argv = ['x', 'none']
exec(open('/root/repo/tools/sweep_ls.py').read().split("what = sys.argv")[0])
BASE.clear(); BASE.update(ls_history=1)
L = [dict(ls_hist_a=a, ls_hist_b=b) for a, b in ((250, 105), (200, 120), (300, 130), (400, 150), (150, 105), (250, 100), (1000000, 105))]
g = ctx.build_torus(650.0, 325.0)
W, H = 1920, 1080
cam = api.vdb_render_camera(W, H, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))
sweep('c2', g, cam, W, H, L)
sweep('c2', g, cam, W, H, [dict(ls_tail=t) for t in (24, 48, 96)], shares=(8,))
g.free()
g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0))
W, H = 3840, 2160
cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))
sweep('c4', g, cam, W, H, L[:5], shares=(8,))
sweep('c4', g, cam, W, H, L[:3])
