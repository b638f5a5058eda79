import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
sys.argv = ['x', 'none']
exec(open('/root/repo/tools/sweep_ls.py').read().split("what = sys.argv")[0])
BASE.update(ls_strip=1, ls_refill=32, ls_order=0, ls_affine=0, ls_tail=64, ls_history=0, ls_hist_a=400, ls_hist_b=200)
L = [dict(ls_history=0, rounds=False), dict(ls_history=1, rounds=False), dict(ls_history=1, ls_hist_a=300, ls_hist_b=150, rounds=False), dict(ls_history=1, ls_hist_a=800, ls_hist_b=300, rounds=False),
     dict(ls_history=1, ls_hist_a=200, ls_hist_b=120, rounds=False), dict(ls_history=0), dict(ls_history=1)]
def check(g, cam, W, H):
    film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
    out = []
    for h in (0, 1, 1, 1):
        ctx.set_tuning(ls_history=h)
        film.zero_()
        ctx.render_levelset(g, cam, sh, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=ctx.ls_opts(uniform_bg=True, rounds=False))
        out.append(film.cpu().numpy().copy())
    print("frames identical with and without history:", all(np.array_equal(out[0], o) for o in out[1:]), flush=True)
g = ctx.build_torus(650.0, 325.0)
W, H = 1920, 1080
cam = api.vdb_render_camera(W, H, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))
check(g, cam, W, H)
sweep('c2', g, cam, W, H, L)
sweep('c2', g, cam, W, H, L, shares=(8,))
g.free()
g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0))
W, H = 3840, 2160
cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))
check(g, cam, W, H)
sweep('c4', g, cam, W, H, L[:5])
sweep('c4', g, cam, W, H, L[:5], shares=(8,))
sweep('c4', g, cam, W, H, L[:2], shares=(4, 2))
