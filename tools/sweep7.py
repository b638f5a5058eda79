import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
sys.argv = ['x', 'none']
exec(open('/root/repo/tools/sweep_ls.py').read().split("what = sys.argv")[0])
BASE.update(ls_strip=1, ls_refill=32, ls_order=0, ls_affine=0, ls_tail=64)
def leaves(*ks):
    d = dict(ls_rounds=len(ks))
    for i, k in enumerate(ks): d['ls_leaves%d' % i] = k
    return d
L = [dict(rounds=False), leaves(8, 128), leaves(8, 32, 128), leaves(4, 16, 64), leaves(4, 12, 36, 108), leaves(6, 24), leaves(8), leaves(16), leaves(4, 16), leaves(2, 6, 18, 54)]
what = os.environ.get("WHAT", "c2 c4").split()
if 'c2' in what:
    g = ctx.build_torus(650.0, 325.0)
    W, H = 1920, 1080
    cam = api.vdb_render_camera(W, H, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))
    sweep('c2', g, cam, W, H, L)
    sweep('c2', g, cam, W, H, L, shares=(8,))
    g.free()
if 'c4' in what:
    g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0))
    W, H = 3840, 2160
    cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))
    sweep('c4', g, cam, W, H, L, shares=(8,))
    if 'prof' in what:
        film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
        timeit(g, cam, W, H, film, part=api.partition(0, 8, 64, 60), n=1, **L[1])
