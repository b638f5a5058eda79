# per-rank kernel time of a partitioned frame, all ranks run one after the other on ONE GPU
# usage: prof_part.py [c2|c4] [n ...]      (environment: VDBRT_LS_BUDGET / VDBRT_LS_FACTOR / VDBRT_LS_ROUNDS)
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from openvdb_b200 import api, _abi as abi
wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
ns = [int(a) for a in sys.argv[2:]] or [2, 4, 8]
ctx = api.Context(0)
if wl == 'c4':
    big = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0)); W, H = 3840, 2160
    cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3.0 * 2048.0), (0, 0, 0))
else:
    big = ctx.build_torus(650.0, 325.0); W, H = 1920, 1080
    cam = api.vdb_render_camera(W, H, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))
sh = api.make_shader(abi.SHADER_DIFFUSE)
film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
def run(o):
    ts = []
    for it in range(3):
        ctx.render_levelset(big, cam, sh, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=o)
        ts.append(ctx.last_kernel_ms()[0])
    return min(ts)
print(wl, 'whole frame: in line %.3f ms, with rounds %.3f ms' % (run(ctx.ls_opts(uniform_bg=True, rounds=False)), run(ctx.ls_opts(uniform_bg=True, rounds=True))))
for n in ns:
    for rounds in (False, True):
        t = [run(ctx.ls_opts(part=api.partition(r, n, 64, 60), uniform_bg=True, rounds=rounds)) for r in range(n)]
        print('n', n, 'rounds' if rounds else 'inline', 'per-rank ms', ' '.join('%.3f' % x for x in t), 'max %.3f' % max(t))
