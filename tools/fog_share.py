# one rank's share of the C3 fog frame (wavefront): kernel time per share; run under `ncu --metrics gpu__time_duration.sum` for the per-kernel list
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from openvdb_b200 import api, _abi as abi
ctx = api.Context(0)
ls = ctx.build_sphere(509.0); fog = ctx.build_fog(ls); ls.free()
W, H = 1920, 1080
cam = api.vdb_render_camera(W, H, (0, 0, 3 * 509.0), (0, 0, 0))
film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
for share in (1, 2, 4, 8):
    per = []
    for r in range(share):
        vo = api.vol_opts_default(); vo.primary_step = 0.5
        if share > 1: vo.part = api.partition(r, share, 64, 60)
        ms = []
        for it in range(3):
            ctx.render_volume(fog, cam, vo, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE); ms.append(ctx.last_kernel_ms()[0])
        per.append(min(ms))
    print("c3 1/%d: max %.3f mean %.3f ms" % (share, max(per), sum(per) / len(per)), flush=True)
