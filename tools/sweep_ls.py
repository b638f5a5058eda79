# times the level-set frame of C2 (and, with `c4`, C4) for a list of scheduler settings in ONE process (vdbrt_set_tuning):
# device-resident film, CUDA-event kernel time incl. the probe launch, median and min of 7 frames after 2 warm-ups.
# usage: tools/sweep_ls.py [c2|c4|c2share8|c4share8]...
import sys, itertools, numpy as np, torch
sys.path.insert(0, '/root/repo')
from openvdb_b200 import api, _abi as abi

ctx = api.Context(0)
sh = api.make_shader(abi.SHADER_DIFFUSE)
BASE = dict(ls_strip=8, ls_strip_ratio=4, ls_refill=16, ls_eager=0, ls_order=2, ls_probe_cap=128, ls_probe_b=64)


def timeit(grid, cam, W, H, film, part=None, n=7, **tune):
    t = dict(BASE); t.update(tune); t.pop('rounds', None)
    ctx.set_tuning(**t)
    opts = ctx.ls_opts(uniform_bg=True, part=part, rounds=tune.get('rounds'))
    ms = []
    for it in range(n + 2):
        ctx.render_levelset(grid, cam, sh, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=opts)
        ms.append(ctx.last_kernel_ms()[0])
    ms = ms[2:]
    return float(np.median(ms)), float(min(ms)), ctx.last_kernel_ms()[1]


def sweep(name, grid, cam, W, H, settings, shares=(1,)):
    film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
    for share in shares:
        for s in settings:
            if share == 1:
                med, mn, nl = timeit(grid, cam, W, H, film, **s)
                print("%-10s %-70s med %7.3f  min %7.3f ms  launches %d" % (name, s, med, mn, nl), flush=True)
            else:
                per = [timeit(grid, cam, W, H, film, part=api.partition(r, share, 64, 60), n=3, **s)[1] for r in range(share)]
                print("%-10s 1/%d %-66s max %7.3f  mean %7.3f ms" % (name, share, s, max(per), sum(per) / len(per)), flush=True)


SET_A = [dict(ls_refill=32, ls_order=0, ls_strip=1),              # round 1's scheduling
         dict(ls_refill=32, ls_order=0),
         dict(ls_refill=32, ls_order=1),
         dict(ls_refill=32, ls_order=1, ls_probe_cap=64, ls_probe_b=32),
         dict(ls_refill=32, ls_order=1, ls_probe_cap=256, ls_probe_b=96),
         dict(ls_refill=24, ls_order=0), dict(ls_refill=16, ls_order=0), dict(ls_refill=8, ls_order=0), dict(ls_refill=4, ls_order=0),
         dict(ls_refill=16, ls_order=0, ls_strip=4), dict(ls_refill=16, ls_order=0, ls_strip=16), dict(ls_refill=16, ls_order=0, ls_strip=32),
         dict(ls_refill=16, ls_order=0, ls_eager=1), dict(ls_refill=8, ls_order=0, ls_eager=1),
         dict(ls_refill=16, ls_order=1), dict(ls_refill=8, ls_order=1), dict(ls_refill=16, ls_order=1, ls_eager=1),
         dict(ls_refill=16, ls_order=1, ls_strip=16), dict(ls_refill=12, ls_order=1, ls_strip=16),
         dict(ls_refill=16, ls_order=1, ls_probe_cap=64, ls_probe_b=32), dict(ls_refill=16, ls_order=1, ls_probe_cap=256, ls_probe_b=96)]
SET_SHARE = [dict(ls_refill=32, ls_order=0, ls_strip=1), dict(ls_refill=32, ls_order=0, ls_strip=1, rounds=False),
             dict(ls_refill=32, ls_order=1, ls_strip=1, rounds=False), dict(ls_refill=32, ls_order=1, ls_strip=1),
             dict(ls_refill=16, ls_order=1, rounds=False), dict(ls_refill=16, ls_order=1, ls_eager=1, rounds=False),
             dict(ls_refill=16, ls_order=1, ls_strip_ratio=2, rounds=False), dict(ls_refill=16, ls_order=1, ls_strip_ratio=1, ls_eager=1, rounds=False)]
what = sys.argv[1:] or ['c2']
if any(w.startswith('c2') for w in what):
    g = ctx.build_torus(650.0, 325.0)
    W, H = 1920, 1080
    cam = api.vdb_render_camera(W, H, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))
    if 'c2' in what: sweep('c2', g, cam, W, H, SET_A)
    if 'c2share8' in what: sweep('c2', g, cam, W, H, SET_SHARE, shares=(8,))
    if 'c2share4' in what: sweep('c2', g, cam, W, H, SET_SHARE, shares=(4,))
    g.free()
if any(w.startswith('c4') for w in what):
    g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0))
    W, H = 3840, 2160
    cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))
    if 'c4' in what: sweep('c4', g, cam, W, H, [SET_A[i] for i in (0, 1, 2, 6, 7, 12, 14, 15, 17)])
    if 'c4share8' in what: sweep('c4', g, cam, W, H, [SET_SHARE[i] for i in (1, 2, 4, 5)], shares=(8,))
    g.free()
