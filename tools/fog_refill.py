import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from openvdb_b200 import api, _abi as abi
ctx = api.Context(0)
def go(name, fog, cam, W, H, spp, part, reps=3):
    film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
    ref = None
    for refill in ((1, 1), (1, 2), (1, 4), (1, 8), (1, 16), (2, 4), (4, 4), (2, 8)):
        ctx.set_tuning(fog_wave=1, fog_refill=8, fog_run_primary=refill[0], fog_run_shadow=refill[1])
        vo = api.vol_opts_default(spp=spp, seed=0)
        vo.primary_step = 0.5
        if part is not None: vo.part = part
        ms = []
        for it in range(reps):
            film.zero_()
            ctx.render_volume(fog, cam, vo, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE)
            ms.append(ctx.last_kernel_ms()[0])
        f = film.cpu().numpy().copy()
        if ref is None: ref = f
        print("%-12s run %s: min %.3f med %.3f ms same %s" % (name, refill, min(ms), float(np.median(ms)), np.array_equal(f, ref)), flush=True)
ls = ctx.build_sphere(509.0); fog = ctx.build_fog(ls); ls.free()
go('c3', fog, api.vdb_render_camera(1920, 1080, (0, 0, 3 * 509.0), (0, 0, 0)), 1920, 1080, 1, None, reps=4)
fog.free()
g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0)); fog = ctx.build_fog(g); g.free()
cam = api.vdb_render_camera(3840, 2160, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))
go('c5 1/64', fog, cam, 3840, 2160, 16, api.partition(0, 64, 64, 60), reps=2)
