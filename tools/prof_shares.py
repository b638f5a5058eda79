# per-rank kernel time of a partitioned c2 frame, every rank's share rendered on ONE GPU: max / mean over ranks per partition tile size
import sys, torch
sys.path.insert(0, '/root/repo')
from openvdb_b200 import api, _abi as abi
ctx = api.Context(0)
big = ctx.build_torus(650.0, 325.0); W, H = 1920, 1080
cam = api.vdb_render_camera(W, H, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))
sh = api.make_shader(abi.SHADER_DIFFUSE)
film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
for count in (2, 4, 8):
    for tw, th in ((64, 60), (32, 60), (32, 30), (64, 20), (16, 20), (128, 120)):
        ms = []
        for r in range(count):
            best = 1e9
            for it in range(3):
                ctx.render_levelset(big, cam, sh, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=ctx.ls_opts(part=api.partition(r, count, tw, th), uniform_bg=True))
                best = min(best, ctx.last_kernel_ms()[0])
            ms.append(best)
        print("ranks %d tile %3dx%-3d  max %.3f  mean %.3f  min %.3f ms" % (count, tw, th, max(ms), sum(ms) / len(ms), min(ms)), flush=True)
