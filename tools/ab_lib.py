# times the C2 and C4 level-set frames (plain order and with the tile history) with whatever library VDBRT_LIBRARY names;
# prints a checksum of the film so that variants can be compared for bit-equality.  usage: tools/ab_lib.py [c2] [c4] [c4share8]
import sys, zlib, numpy as np, torch
sys.path.insert(0, '/root/repo')
from openvdb_b200 import api, _abi as abi

what = sys.argv[1:] or ['c2', 'c4']
ctx = api.Context(0)
sh = api.make_shader(abi.SHADER_DIFFUSE)


def frame(grid, cam, W, H, hist, part=None, n=7):
    ctx.set_tuning(ls_history=hist)
    film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
    opts = ctx.ls_opts(uniform_bg=True, part=part)
    ms = []
    for it in range(n + 3):
        ctx.render_levelset(grid, cam, sh, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=opts)
        ms.append(ctx.last_kernel_ms()[0])
    return float(np.median(ms[3:])), float(min(ms[3:])), zlib.crc32(film.cpu().numpy().tobytes())


def run(name, grid, cam, W, H):
    for hist in (0, 1):
        if name in what:
            med, mn, crc = frame(grid, cam, W, H, hist)
            print("%-4s history %d  med %7.3f  min %7.3f ms  crc %08x" % (name, hist, med, mn, crc), flush=True)
        if name + 'share8' in what:
            per = [frame(grid, cam, W, H, hist, part=api.partition(r, 8, 64, 60), n=3)[1] for r in range(8)]
            print("%-4s 1/8 history %d  max %7.3f  mean %7.3f ms" % (name, hist, max(per), sum(per) / 8), flush=True)


if any(w.startswith('c2') for w in what):
    g = ctx.build_torus(650.0, 325.0)
    run('c2', g, api.vdb_render_camera(1920, 1080, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0)), 1920, 1080)
    g.free()
if any(w.startswith('c4') for w in what):
    g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0))
    run('c4', g, api.vdb_render_camera(3840, 2160, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0)), 3840, 2160)
    g.free()
