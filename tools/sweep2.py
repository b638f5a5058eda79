import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
sys.argv = ['x', 'none']
exec(open('/root/repo/tools/sweep_ls.py').read().split("what = sys.argv")[0])
g = ctx.build_torus(650.0, 325.0)
W, H = 1920, 1080
cam = api.vdb_render_camera(W, H, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))
S = [dict(ls_refill=r, ls_order=0, ls_strip=s, ls_eager=e) for s in (1, 2, 3) for r in (32, 16, 8) for e in (0, 1) if not (s == 1 and e == 0 and r != 32)]
sweep('c2', g, cam, W, H, S)
S = [dict(ls_refill=32, ls_order=1, ls_strip=1, ls_probe_cap=c, ls_probe_b=c // 2) for c in (8, 16, 32, 64, 128)]
sweep('c2', g, cam, W, H, S)
sweep('c2', g, cam, W, H, [dict(ls_refill=32, ls_order=0, ls_strip=1, rounds=False)] + [dict(rounds=False, **s) for s in S], shares=(8,))
