# A/B of the fog paths (fog_wave 0 = one-loop kernel, 1 = wavefront): C3 (1080p) and a share of C5's fog pass (4K, 16 spp)
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from openvdb_b200 import api, _abi as abi
ctx = api.Context(0)
def go(name, fog, cam, W, H, spp, part, reps=3):
    film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
    frames = []
    for wave in (0, 1):
        ctx.set_tuning(fog_wave=wave)
        vo = api.vol_opts_default(spp=spp, seed=0)
        vo.primary_step = 0.5
        if part is not None: vo.part = part
        ms = []
        for it in range(reps):
            film.zero_()
            ctx.render_volume(fog, cam, vo, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE)
            ms.append(ctx.last_kernel_ms()[0])
        frames.append(film.cpu().numpy().copy())
        print("%-12s wave %d: min %.3f med %.3f ms, launches %d" % (name, wave, min(ms), float(np.median(ms)), ctx.last_kernel_ms()[1]), flush=True)
    print("%-12s frames identical: %s" % (name, np.array_equal(frames[0], frames[1])), flush=True)
what = sys.argv[1:] or ['c3', 'c5']
if 'c3' in what:
    ls = ctx.build_sphere(509.0); fog = ctx.build_fog(ls); ls.free()
    go('c3', fog, api.vdb_render_camera(1920, 1080, (0, 0, 3 * 509.0), (0, 0, 0)), 1920, 1080, 1, None, reps=5)
    fog.free()
if 'c5' in what:
    g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0)); fog = ctx.build_fog(g); g.free()
    cam = api.vdb_render_camera(3840, 2160, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))
    go('c5 1/64', fog, cam, 3840, 2160, 16, api.partition(0, 64, 64, 60), reps=2)
    if "c5full" in what: go("c5 1/8", fog, cam, 3840, 2160, 16, api.partition(0, 8, 64, 60), reps=2)
