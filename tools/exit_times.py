# VDBRT_DEBUG_EXIT=1: when do the warps of the render kernel leave?  c2 and c4, whole frame and a 1/8 share
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from openvdb_b200 import api, _abi as abi
ctx = api.Context(0)
sh = api.make_shader(abi.SHADER_DIFFUSE)
def go(name, g, cam, W, H):
    film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
    for part in (None, api.partition(0, 8, 64, 60), api.partition(3, 8, 64, 60)):
        for it in range(3):
            print(name, "whole" if part is None else "rank %d of 8" % part.rank, flush=True, file=sys.stderr)
            ctx.render_levelset(g, cam, sh, film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=ctx.ls_opts(uniform_bg=True, part=part, rounds=False))
g = ctx.build_torus(650.0, 325.0)
go('c2', g, api.vdb_render_camera(1920, 1080, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0)), 1920, 1080)
g.free()
g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0))
go('c4', g, api.vdb_render_camera(3840, 2160, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0)), 3840, 2160)
