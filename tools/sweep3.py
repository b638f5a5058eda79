import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
from openvdb_b200 import api, _abi as abi
ctx = api.Context(0)
sh = api.make_shader(abi.SHADER_DIFFUSE)
g = ctx.build_torus(650.0, 325.0)
W, H = 1920, 1080
cam = api.vdb_render_camera(W, H, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))
film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
ctx.set_tuning(ls_strip=1, ls_refill=32, ls_order=0)
ref = None
for budget, factor, rounds in ((0, 50, False), (160, 50, True), (64, 0, True), (16, 0, True), (1, 0, True)):
    if budget: ctx.set_tuning(ls_budget=budget, ls_factor=factor)
    opts = ctx.ls_opts(uniform_bg=True, rounds=rounds)
    ms = []
    for it in range(6):
        ctx.render_levelset(g, cam, sh, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=opts)
        ms.append(ctx.last_kernel_ms()[0])
    f = film.cpu().numpy()
    if ref is None: ref = f
    print("budget %d factor %d rounds %s leaves %s: med %.3f min %.3f ms launches %d same %s" % (budget, factor, rounds, os.environ.get("VDBRT_LS_LEAVES"), np.median(ms[1:]), min(ms), ctx.last_kernel_ms()[1], np.array_equal(f, ref)), flush=True)
