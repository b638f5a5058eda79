import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
sys.argv = ['x', 'none']
exec(open('/root/repo/tools/sweep_ls.py').read().split("what = sys.argv")[0])
BASE.update(ls_strip=1, ls_refill=32, ls_order=0, ls_affine=0)
g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0))
W, H = 3840, 2160
cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))
S = [dict(ls_order=0), dict(ls_order=0, ls_strip=2), dict(ls_order=1, ls_probe_cap=64, ls_probe_b=32), dict(ls_order=1, ls_probe_cap=128, ls_probe_b=64), dict(ls_order=1, ls_probe_cap=256, ls_probe_b=128),
     dict(ls_order=0, ls_budget=160, ls_factor=50, rounds=True)]
sweep('c4', g, cam, W, H, S)
S = [dict(ls_order=0, rounds=False), dict(ls_order=1, ls_probe_cap=64, ls_probe_b=32, rounds=False), dict(ls_order=1, ls_probe_cap=128, ls_probe_b=64, rounds=False),
     dict(ls_order=1, ls_probe_cap=256, ls_probe_b=128, rounds=False), dict(ls_order=0, rounds=True)]
sweep('c4', g, cam, W, H, S, shares=(8,))
sweep('c4', g, cam, W, H, S[:3], shares=(2,))
