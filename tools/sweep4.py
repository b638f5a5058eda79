import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
sys.argv = ['x', 'none']
exec(open('/root/repo/tools/sweep_ls.py').read().split("what = sys.argv")[0])
BASE.update(ls_strip=1, ls_refill=32, ls_order=0, ls_affine=0)
g = ctx.build_torus(650.0, 325.0)
W, H = 1920, 1080
cam = api.vdb_render_camera(W, H, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))
film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
ctx.set_tuning(**BASE)
ctx.render_levelset(g, cam, sh, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=ctx.ls_opts(uniform_bg=True))
ref = film.cpu().numpy().copy()
S = [dict(ls_affine=a) for a in (0, 8, 16, 32, 64, 128, 256, 512)]
sweep('c2', g, cam, W, H, S)
ctx.set_tuning(ls_affine=128)
film.zero_()
ctx.render_levelset(g, cam, sh, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=ctx.ls_opts(uniform_bg=True))
print("affine frame identical:", np.array_equal(ref, film.cpu().numpy()))
sweep('c2', g, cam, W, H, [dict(ls_affine=a, rounds=r) for a in (0, 16, 128) for r in (False, True)], shares=(8,))
