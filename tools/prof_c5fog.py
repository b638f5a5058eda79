# one small share of the C5 fog pass (the C4 union's fog volume, 3840x2160, 16 samples per pixel): rank 0 of an N-way split
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from openvdb_b200 import api, _abi as abi
ctx = api.Context(0)
share = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0))
fog = ctx.build_fog(g)
g.free()
W, H = 3840, 2160
cam = api.vdb_render_camera(W, H, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))
film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
vo = api.vol_opts_default(spp=16, seed=0)
vo.primary_step = 0.5
vo.part = api.partition(0, share, 64, 60)
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    ctx.render_volume(fog, cam, vo, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE)
    print(ctx.last_kernel_ms(), flush=True)
if len(sys.argv) > 3:
    vo1 = api.vol_opts_default(); vo1.primary_step = 0.5; vo1.part = api.partition(0, share, 64, 60)
    print(ctx.count_volume(fog, cam, vo1).as_dict())
