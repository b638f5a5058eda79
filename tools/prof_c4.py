# renders the c4 frame (union of 10 000 spheres, 3840x2160) N times with a device-resident film: `ncu -k regex:k_render_levelset -s 2 -c 1`
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from openvdb_b200 import api, _abi as abi
ctx = api.Context(0)
g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0))
W, H = 3840, 2160
cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))
sh = api.make_shader(abi.SHADER_DIFFUSE)
film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ctx.render_levelset(g, cam, sh, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=ctx.ls_opts(uniform_bg=True, rounds=False))
    print(ctx.last_kernel_ms())
if len(sys.argv) > 2:
    print(ctx.count_levelset(g, cam).as_dict())
