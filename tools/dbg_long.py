import sys, torch
sys.path.insert(0, '/root/repo')
from openvdb_b200 import api, _abi as abi
ctx = api.Context(0)
wl=sys.argv[1]
if wl == 'c4':
    big = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0)); W, H = 3840, 2160
    cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3.0 * 2048.0), (0, 0, 0))
else:
    big = ctx.build_torus(650.0, 325.0); W, H = 1920, 1080
    cam = api.vdb_render_camera(W, H, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))
sh = api.make_shader(abi.SHADER_DIFFUSE)
film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
for part in (None, api.partition(0,2,64,60), api.partition(0, 8, 64, 60)):
    ctx.render_levelset(big, cam, sh, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=ctx.ls_opts(part=part, uniform_bg=True, rounds=True))
    print(ctx.last_kernel_ms())
