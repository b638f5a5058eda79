import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
sys.argv = ['x', 'none']
exec(open('/root/repo/tools/sweep_ls.py').read().split("what = sys.argv")[0])
BASE.clear(); BASE.update(ls_history=1)
L = [dict()]
def once(name, g, cam, W, H, part=None):
    film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
    print(name, flush=True)
    timeit(g, cam, W, H, film, part=part, n=1)
g = ctx.build_torus(650.0, 325.0)
cam = api.vdb_render_camera(1920, 1080, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))
once('c2', g, cam, 1920, 1080); once('c2 1/2', g, cam, 1920, 1080, api.partition(0, 2, 64, 60))
g.free()
g = ctx.build_sphere(100.0)
once('c1', g, api.vdb_render_camera(1024, 1024, (0, 0, 300.0), (0, 0, 0)), 1024, 1024)
g.free()
g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0))
cam = api.vdb_render_camera(3840, 2160, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))
once('c4', g, cam, 3840, 2160)
for s in (2, 4, 8): once('c4 1/%d' % s, g, cam, 3840, 2160, api.partition(0, s, 64, 60))
