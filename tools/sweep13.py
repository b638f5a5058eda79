import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
sys.argv = ['x', 'none']
exec(open('/root/repo/tools/sweep_ls.py').read().split("what = sys.argv")[0])
BASE.clear(); BASE.update(ls_history=1)
g = ctx.build_torus(650.0, 325.0)
W, H = 1920, 1080
cam = api.vdb_render_camera(W, H, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))
sweep('c2', g, cam, W, H, [dict(ls_tail=t) for t in (8, 12, 16, 24, 32)], shares=(8,))
sweep('c2', g, cam, W, H, [dict(ls_tail=t) for t in (16, 24)], shares=(4, 2))
g.free()
