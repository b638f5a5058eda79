import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from openvdb_b200 import api, _abi as abi
ctx=api.Context(0)
big=ctx.build_torus(650.0,325.0)
W,H=1920,1080
cam=api.vdb_render_camera(W,H,(0,1.5*650,3*(650+325.0)),(0,0,0))
sh=api.make_shader(abi.SHADER_DIFFUSE)
film = torch.zeros((H, W, 4), dtype=torch.float32, device='cuda')
for it in range(int(sys.argv[1]) if len(sys.argv)>1 else 3):
    ctx.render_levelset(big,cam,sh,film.data_ptr(), width=W, height=H, memspace=abi.MEM_DEVICE, opts=ctx.ls_opts(uniform_bg=True, rounds=True))
    print(ctx.last_kernel_ms())
