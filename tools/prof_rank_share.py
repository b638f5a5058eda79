# one rank's share (rank 0 of 8) of the c2 frame, rendered 4 times: the launch list shows the long-ray round kernels
import sys, torch
sys.path.insert(0, '/root/repo')
from openvdb_b200 import api, _abi as abi
ctx = api.Context(0)
big = ctx.build_torus(650.0, 325.0); W, H = 1920, 1080
cam = api.vdb_render_camera(W, H, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))
sh = api.make_shader(abi.SHADER_DIFFUSE)
film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
for it in range(4):
    ctx.render_levelset(big, cam, sh, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=ctx.ls_opts(part=api.partition(0, 8, 64, 60), uniform_bg=True))
    print(ctx.last_kernel_ms())
