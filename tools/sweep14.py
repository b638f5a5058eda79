import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
sys.argv = ['x', 'none']
exec(open('/root/repo/tools/sweep_ls.py').read().split("what = sys.argv")[0])
BASE.clear(); BASE.update(ls_history=1)
g = ctx.build_torus(650.0, 325.0)
W, H = 1920, 1080
cam = api.vdb_render_camera(W, H, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))
L = [dict(ls_rounds=1, ls_leaves0=k) for k in (40, 48, 64, 80, 96)] + [dict(ls_rounds=2, ls_leaves0=64, ls_leaves1=256), dict(ls_rounds=2, ls_leaves0=48, ls_leaves1=128)]
sweep('c2', g, cam, W, H, L, shares=(8,))
sweep('c2', g, cam, W, H, L, shares=(4,))
g.free()
