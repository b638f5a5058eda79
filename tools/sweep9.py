import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
sys.argv = ['x', 'none']
exec(open('/root/repo/tools/sweep_ls.py').read().split("what = sys.argv")[0])
BASE.update(ls_strip=1, ls_refill=32, ls_order=0, ls_affine=0, ls_tail=64, ls_voxel_only=0, ls_run_max=0, ls_run_min=8)
L = [dict(rounds=False, ls_run_max=0)] + [dict(rounds=False, ls_run_max=m, ls_run_min=n) for m, n in ((4, 8), (8, 8), (16, 8), (8, 4), (8, 12), (8, 16), (16, 16), (32, 8), (16, 4), (8, 1))]
def check(g, cam, W, H):
    film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
    out = []
    for rm in (0, 8):
        ctx.set_tuning(ls_run_max=rm, ls_run_min=2)
        film.zero_()
        ctx.render_levelset(g, cam, sh, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=ctx.ls_opts(uniform_bg=True, rounds=False))
        out.append(film.cpu().numpy().copy())
    print("frames identical with and without the run:", np.array_equal(out[0], out[1]), flush=True)
g = ctx.build_torus(650.0, 325.0)
W, H = 1920, 1080
cam = api.vdb_render_camera(W, H, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))
check(g, cam, W, H)
sweep('c2', g, cam, W, H, L)
g.free()
g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0))
W, H = 3840, 2160
cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))
check(g, cam, W, H)
sweep('c4', g, cam, W, H, L)
sweep('c4', g, cam, W, H, L[:3], shares=(8,))
