# rank 0's share (1/8) of the C5 fog pass (4K, 16 spp, the 10k-sphere union's fog), three times: first call (buffers are allocated) and warm
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from openvdb_b200 import api, _abi as abi
ctx = api.Context(0)
g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0)); fog = ctx.build_fog(g)
W, H = 3840, 2160
cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))
part = api.partition(0, 8, 64, 60)
film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
for it in range(3):
    ctx.render_levelset(g, cam, api.make_shader(abi.SHADER_DIFFUSE), film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE,
                        opts=ctx.ls_opts(spp=16, seed=0, uniform_bg=True, part=part))
    ms_ls = ctx.last_kernel_ms()
    vo = api.vol_opts_default(spp=16, seed=0); vo.primary_step = 0.5; vo.part = part
    ctx.render_volume(fog, cam, vo, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE)
    print("c5 1/8 share, call %d: level set %.1f ms (%d launches), fog %.1f ms (%d launches)" % ((it,) + tuple(ms_ls) + tuple(ctx.last_kernel_ms())), flush=True)
